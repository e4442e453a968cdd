// le_inst_cp_tanh.cu — compiled kernel set for SD=4, AD=2, QACT_TANH, units per thread {2,4,6} (Q-net hidden <= 32*U: up to 192, the DDQN_vary range [19,171]).
#include "le_instance.cuh"
namespace le {
void le_register_cp_tanh() {
    le_register_instance(InstanceImpl<4, 2, 2, QACT_TANH>::ops());
    le_register_instance(InstanceImpl<4, 2, 4, QACT_TANH>::ops());
    le_register_instance(InstanceImpl<4, 2, 6, QACT_TANH>::ops());
}
}  // namespace le
