// placeholder until the tcgen05 kernel lands (replaced in the next commit)
#include "sa_common.cuh"
extern "C" int sa_corr_tf32(const float*, const float*, float*, int, int, int, int, int, float, float, const float*,
                            const float*, double, float*, float*, float*, int64_t, int64_t, int64_t, void*) {
  SA_FAIL(SA_E_UNSUPPORTED, "sa_corr_tf32: not built yet");
}
