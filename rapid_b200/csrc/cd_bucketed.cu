// placeholder until the subject-bucketed kernels land
#include "cd_internal.cuh"
namespace rapid {
int32_t bucketed_apply(CD*, int64_t, const DeliveryDev&, const BatchCounts&) { set_error("bucketed kernels not built"); return RAPID_EUNSUPPORTED; }
void bucketed_destroy(CD*) {}
int32_t bucketed_clear(CD*) { return RAPID_OK; }
}
