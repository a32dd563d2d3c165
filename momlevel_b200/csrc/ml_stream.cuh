// ml_stream.cuh -- interface of the ring-staged elementwise kernels (ml_stream.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ml {
namespace stream {

// rows of ncol fp32 points start 16-byte aligned in every operand (pass NULL for an absent one)
bool eligible(const void* a, const void* b, const void* c, const void* out, long long ncol);

// spice.flament.spice over n points (n % 4 == 0)
int launch_spice(const float* T, const float* S, long long n, double* out, cudaStream_t st);

// eos.<name>.density over rows [nouter * nz][ncol]; p scalar / per level / absent (pmode as ml_eos_eval)
int launch_density(int eos, const float* T, const float* S, long long t_stride, long long s_stride, const double* p,
                   int pmode, long long nrows, int nz, long long ncol, double* out, cudaStream_t st);

// reference-state pass: rho_ref out, block partials [2][refstate_blocks()] for the fixed-order second stage
int refstate_blocks(long long nz, long long ncol);
int launch_refstate(int eos, const float* T0, const float* S0, const float* V0, const double* p_level, long long nz,
                    long long ncol, double* rho_ref, double* partials, cudaStream_t st);

// delta_rho[t][z][col] = rho(T, S, p_z) - rho_ref where the reference volume is present, NaN elsewhere (steric.py:151-153);
// t_stride / s_stride = 0 for an operand held at the reference slab
int launch_delta_rho(int eos, const float* T, const float* S, long long t_stride, long long s_stride, const double* rho_ref,
                     const void* v_ref, int v_f32, const double* p_level, int nt, long long nz, long long ncol, double* out,
                     cudaStream_t st);

}  // namespace stream
}  // namespace ml
