// le_inst_ac_tanh.cu — compiled kernel set for SD=6, AD=3, QACT_TANH, units per thread {2,4} (Q-net hidden <= 32*U).
#include "le_instance.cuh"
namespace le {
void le_register_ac_tanh() {
    le_register_instance(InstanceImpl<6, 3, 2, QACT_TANH>::ops());
    le_register_instance(InstanceImpl<6, 3, 4, QACT_TANH>::ops());
}
}  // namespace le
