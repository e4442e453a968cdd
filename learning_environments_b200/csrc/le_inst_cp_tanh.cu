// le_inst_cp_tanh.cu — compiled kernel set for SD=4, AD=2, QACT_TANH, units per thread {2,4} (Q-net hidden <= 32*U).
#include "le_instance.cuh"
namespace le {
void le_register_cp_tanh() {
    le_register_instance(InstanceImpl<4, 2, 2, QACT_TANH>::ops());
    le_register_instance(InstanceImpl<4, 2, 4, QACT_TANH>::ops());
}
}  // namespace le
