// le_inst_ac_leaky.cu — compiled kernel set for SD=6, AD=3, QACT_LEAKY, units per thread {2,4} (Q-net hidden <= 32*U).
#include "le_instance.cuh"
namespace le {
void le_register_ac_leaky() {
    le_register_instance(InstanceImpl<6, 3, 2, QACT_LEAKY>::ops());
    le_register_instance(InstanceImpl<6, 3, 4, QACT_LEAKY>::ops());
}
}  // namespace le
