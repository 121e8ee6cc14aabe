#include "common.cuh"
namespace sbce {
cudaError_t launch_pm_stats(const Dims& d, int nb, const double* Yd, const double* PsiD, const double* theta,
                            const double* varn, const int32_t* active, double* stat_m, double* stat_R,
                            cudaStream_t s) {
    return cudaErrorNotSupported;
}
}
