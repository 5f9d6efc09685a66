// placeholder - replaced by the fused kernel
#include "cspn_common.cuh"
namespace cspn {
bool fused_supported(int, int, int, int, int, int) { return false; }
template <typename T> int fused_forward(const FwdArgs<T>&) { return CSPN_ERR_BAD_KERNEL_SIZE; }
template int fused_forward<float>(const FwdArgs<float>&);
template int fused_forward<__half>(const FwdArgs<__half>&);
}
