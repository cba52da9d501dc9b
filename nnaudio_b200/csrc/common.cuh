// Shared declarations for libnnab.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nnab.h"

namespace nnab {

// ---- error plumbing -------------------------------------------------------
void set_cuda_error(const char* where, cudaError_t e);
void set_error_text(const char* text);
void count_launch();

#define NNAB_CUDA_TRY(expr)                         \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) {                        \
      ::nnab::set_cuda_error(#expr, _e);            \
      return NNAB_ECUDA;                            \
    }                                               \
  } while (0)

#define NNAB_LAUNCH_CHECK()                         \
  do {                                              \
    ::nnab::count_launch();                         \
    cudaError_t _e = cudaGetLastError();            \
    if (_e != cudaSuccess) {                        \
      ::nnab::set_cuda_error("kernel launch", _e);  \
      return NNAB_ECUDA;                            \
    }                                               \
  } while (0)

// ---- problem description shared by the SIMT and tcgen05 framed kernels ----
// One "framed complex contraction":
//   re[b,f,t] = sum_k xpad[b, t*hop + k] * w_re[f,k]
//   im[b,f,t] = -sum_k xpad[b, t*hop + k] * w_im[f,k]
// followed by a per-bin scale and one of the output formats of nnab.h (or the
// internal POWER format used by the filterbank ops).
constexpr int FMT_POWER = 100;  // internal: (sqrt(re^2+im^2+eps)) ** power -> (B,F,T)
constexpr int FMT_FBANK = 101;  // internal (tcgen05 only): power -> banded filterbank -> (B,n_fb,T)

// Banded filterbank table: the (at most two) non-zero weights of every FFT bin.
struct FbEntry {
  int j0, j1;    // filter rows (-1 = none)
  float w0, w1;
};

struct FramedProblem {
  const float* x;      // (B, L) rows, pitch x_pitch
  int64_t B, L, x_pitch;
  const float* w_re;   // (F, K)
  const float* w_im;   // (F, K)
  int F, K, hop;
  int pad;             // samples of centre padding on each side (0 if !center)
  int pad_mode;        // NNAB_PAD_*
  const float* scale;  // per-bin or nullptr
  float scale_all;
  int fmt;
  float eps;
  float power;         // FMT_POWER only
  float* out;
  int64_t T;
  int out_bins;        // rows of the output tensor (>= bins written)
  int bin_offset;      // output row of bin 0 (may be negative: rows < 0 dropped)
  const int32_t* h_k_begin;  // host, per-bin support or nullptr
  const int32_t* h_k_end;
  const FbEntry* fb_table;   // FMT_FBANK: device table [F]; out is (B, n_fb, T), pre-zeroed
  int n_fb;
};

int launch_framed_simt(const FramedProblem& p, cudaStream_t stream);

// tcgen05 path (tc_kernels.cu)
bool tc_supported(const FramedProblem& p);
size_t tc_workspace_bytes(int64_t B, int64_t L, int K, int hop, int pad);
int launch_framed_tc(const FramedProblem& p, const void* packed, void* workspace,
                     size_t ws_bytes, cudaStream_t stream);
size_t tc_packed_bytes(int F, int K);
int tc_pack_basis(const float* w_re, const float* w_im, int F, int K, void* packed,
                  cudaStream_t stream);
int tc_tile_n();
int launch_fb_table(const float* fb, int n_fb, int F, FbEntry* table, int* d_max_nnz,
                    cudaStream_t stream);

// filterbank / MFCC tail / FIR decimation (simt_kernels.cu)
int launch_filterbank(const float* P, const float* fb, int64_t B, int F, int64_t T, int n_fb,
                      float* out, cudaStream_t stream);
int launch_mfcc_tail(const float* mel, int64_t B, int n_mels, int64_t T, float amin, float ref,
                     float top_db, const float* dct, int n_mfcc, float* out,
                     unsigned int* scratch /* B words */, cudaStream_t stream);
int launch_fir_decimate(const float* x, int64_t B, int64_t L, int64_t x_pitch, const float* fir,
                        int taps, int factor, float* y, int64_t Ly, int64_t y_pitch,
                        cudaStream_t stream);

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nnab
