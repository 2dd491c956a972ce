// tcgen05 engine (placeholder until the UMMA kernels land): reports "unsupported".
#include "model.cuh"

namespace cf {
struct TcEngine {};
bool tc_supported(const HostModel&) { return false; }
TcEngine* tc_create(const HostModel&) { return nullptr; }
void tc_destroy(TcEngine*) {}
int tc_forward(TcEngine*, const HostModel&, const int16_t*, const double*, const float*, WindowTable,
               int64_t, float*, cudaStream_t, Profiler*) {
    set_error("tcgen05 engine not built");
    return CF_ERR_BAD_ARG;
}
}  // namespace cf
