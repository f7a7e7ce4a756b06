// placeholder until the tcgen05 path lands
#include "wn_tc.cuh"
namespace mbx {
void wn_tc_carve(const mbexwn_config_t&, long long, int, const std::function<void(const char*, size_t)>&) {}
int wn_tc_forward(WnTcState&, const mbexwn_config_t&, const FrameGrid&, int, const float*, const float*, float*,
                  const std::function<void*(const char*)>&, const std::function<const void*(const std::string&, size_t)>&,
                  cudaStream_t, int*, std::string* error) {
    if (error) *error = "tensor-core WaveNet path not built";
    return MBEXWN_ERR_UNSUPPORTED;
}
void wn_tc_invalidate(WnTcState&) {}
void wn_tc_destroy(WnTcState&) {}
}  // namespace mbx
