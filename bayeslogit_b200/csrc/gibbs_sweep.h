// Host side of the one-pass logit sweep kernel (gibbs_sweep.cu): psi, omega and X' Omega X from one
// TMA-staged read of X per iteration.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "engine.h"

namespace bl {

struct PeerPush;

// even P <= 64, 16-byte aligned tX (TMA: 16-byte global strides); BL_GIBBS_NO_K3 in the environment disables it
bool logit_sweep_ok(const double *tX, int P);

struct LogitSweep {
    alignas(64) unsigned char map_storage[128];     // CUtensorMap of tX (N x P row-major, 32 x 16 boxes, 128-byte swizzle)
    int64_t N = 0;
    int P = 0, ntiles = 0, grid = 0, draw_warps = 0, gram_warps = 8, stages = 0;
    size_t smem = 0;
    double *part = nullptr;                          // [grid][64 x 64] per-CTA partial tiles
    unsigned *ctr = nullptr;                         // grid barrier counter
    uint64_t launches = 0;
    uint64_t ctr_rounds = 0;                         // increments the grid counter has seen (barrier + publish rounds)
    cudaStream_t stream = nullptr;

    int init(const double *tX, int64_t N, int P, cudaStream_t st, std::string &err);
    // w_out[i] = PG(shape_i, x_i . beta) for the local rows, PP = sum_i w_i x_i x_i' (P x P, no prior);
    // px.world > 1: PP is also stored into the peers' windows (the beta draw then waits for the world's flags)
    cudaError_t launch(double *w_out, double *PP, const int *shape, const double *beta, StreamId id, const PeerPush &px);
    ~LogitSweep();
};

}  // namespace bl
