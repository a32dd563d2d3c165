// ml_tma.cu -- TMA-staged kernel family (under construction: eligibility is off, so every
// launch takes the direct family in ml_api.cu).
#include "ml_tma.cuh"

namespace ml {
namespace tma {

bool local_eligible(int, const void*, const void*, int, int, const double*, const void*, int, int64_t, int64_t, int64_t,
                    const double*, const double*) { return false; }
int launch_local(int, int, const void*, const void*, int, int, const double*, const void*, int, const double*,
                 const double*, const double*, double, int, int, int64_t, double*, double*, cudaStream_t) { return -100; }
bool global_eligible(int, const void*, const void*, int, int, const void*, int, int64_t, int64_t, int64_t) { return false; }
int launch_global(int, int, const void*, const void*, int, int, const void*, int, const double*, int, int, int64_t,
                  double*, double*, cudaStream_t) { return -100; }
bool spice_eligible(int, const void*, const void*, int64_t, const double*) { return false; }
int launch_spice(int, const void*, const void*, int64_t, double*, cudaStream_t) { return -100; }

}  // namespace tma
}  // namespace ml
