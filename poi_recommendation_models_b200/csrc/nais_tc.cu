// placeholder until the tcgen05 path lands
#include "nais_common.cuh"
namespace nais {
bool tc_supported(const NaisParams&) { return false; }
size_t fullrank_tc_workspace_bytes(const NaisParams&, int, int64_t, int64_t, int64_t, int, int) { return 0; }
int launch_fullrank_tc(const NaisParams&, const NaisCatalog&, const NaisUsers&, int64_t, int64_t, int, int, int, float*,
                       int32_t*, float*, void*, size_t, cudaStream_t) { return NAIS_ERR_MODE; }
}  // namespace nais
