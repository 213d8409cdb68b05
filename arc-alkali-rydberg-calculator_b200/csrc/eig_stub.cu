// TEMPORARY: eigensolver entry points not built yet (replaced by eig.cu).
#include "common.cuh"
#define NOTYET(name) { arcb200_set_error(name ": not implemented yet"); return -99; }
extern "C" int64_t arcb200_sweep_workspace_bytes(int32_t, int32_t, int32_t) { return -99; }
extern "C" int arcb200_stark_sweep(int, void *, int32_t, const double *, const double *, int32_t, const double *, int32_t, const double *, int32_t, double, double *, double *, double *, int32_t *, void *, int64_t, int32_t) NOTYET("stark_sweep")
extern "C" int arcb200_pair_sweep(int, void *, int32_t, const double *, int32_t, const double *const *, int32_t, const double *, int32_t, double, int32_t, const double *, int32_t, double *, double *, double *, int32_t *, void *, int64_t, int32_t) NOTYET("pair_sweep")
extern "C" int arcb200_scatter_coo(int, void *, int32_t, int64_t, const int32_t *, const int32_t *, const double *, double *) NOTYET("scatter_coo")
extern "C" int arcb200_syevd_batched(int, void *, int32_t, int32_t, double *, double *, double *, void *, int64_t) NOTYET("syevd")
