// Pipe-throughput microbenchmarks: the denominators of the sampler's compute roofline.
//
// MEASURED_PEAKS.json (driver-written) holds the HBM copy bandwidth and the cuBLAS bf16 rate; the
// sampler kernels are bound by neither (16-24 bytes per draw), so bench.py
// measures, in the same run and at the clocks of that run, what the SMs sustain on the pipes the
// sampler uses: FP64 FMA, FP32 FMA, MUFU (ex2/lg2), the FP64 tensor path (DMMA m8n8k4), 32-bit integer
// multiply-add (Philox rounds), and the warp-instruction issue rate (independent IADD3s, which is what
// "issue slots" means in ncu's sm__inst_issued).  Each kernel keeps 8 independent dependency chains per
// thread, 1024 threads per CTA, 2 CTAs per SM: enough to cover every pipe's latency.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "engine.h"

namespace bl {

namespace {

constexpr int kIters = 2048;
constexpr int kChains = 8;

__global__ void __launch_bounds__(1024) k_peak_dfma(double *out, double a, double b)
{
    double x[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(1024) k_peak_ffma(float *out, float a, float b)
{
    float x[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = threadIdx.x * 1e-3f + k;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) x[k] = fmaf(x[k], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k];
    if (s == 12345.678f) out[0] = s;
}

// ex2 followed by lg2: two MUFU operations per pair, the value stays put
__global__ void __launch_bounds__(1024) k_peak_mufu(float *out)
{
    float x[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = 0.5f + threadIdx.x * 1e-4f + 0.01f * k;
    for (int it = 0; it < kIters / 2; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) {
            float y;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[k]));
            asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(x[k]) : "f"(y));
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k];
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(1024) k_peak_dmma(double *out, double a, double b)
{
    double c[kChains][2];
#pragma unroll
    for (int k = 0; k < kChains; ++k) { c[k][0] = k; c[k][1] = -k; }
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += c[k][0] + c[k][1];
    if (s == 12345.678) out[0] = s;
}

// The same with operands that change from one MMA to the next, as in a real contraction: every MMA reads its
// own A and B registers (here: rotated through eight live values), so nothing comes from the operand reuse cache.
__global__ void __launch_bounds__(1024) k_peak_dmma_operands(double *out, double a, double b)
{
    double c[kChains][2], av[kChains], bv[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) { c[k][0] = k; c[k][1] = -k; av[k] = a + 1e-3 * (threadIdx.x + k); bv[k] = b + 1e-3 * (threadIdx.x ^ k); }
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(av[k]), "d"(bv[(k + it) & (kChains - 1)]));
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += c[k][0] + c[k][1];
    if (s == 12345.678) out[0] = s;
}

// ... and with the weighted-Gram pattern: the A operand of every group of MMAs is the product of a fragment
// and a weight, formed by a DMUL right in front of them (kMulEvery MMAs per DMUL).  DMUL and DMMA share the
// FP64 pipe; this measures what the interleaving costs.
template <int kMulEvery>
__global__ void __launch_bounds__(1024) k_peak_dmma_mul(double *out, double a, double b)
{
    double c[kChains][2], av[kChains], bv[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) { c[k][0] = k; c[k][1] = -k; av[k] = a + 1e-3 * (threadIdx.x + k); bv[k] = b + 1e-3 * (threadIdx.x ^ k); }
    double w = 1.0 + 1e-9 * threadIdx.x;
    for (int it = 0; it < kIters; ++it) {
        double aw = 0.0;
#pragma unroll
        for (int k = 0; k < kChains; ++k) {
            if (k % kMulEvery == 0) aw = av[k] * w;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(aw), "d"(bv[(k + it) & (kChains - 1)]));
        }
        w += 1e-12;
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += c[k][0] + c[k][1];
    if (s == 12345.678) out[0] = s;
}

// the Philox round's multiply: 32 x 32 -> 64 bit (IMAD.WIDE.U32), folded back to 32 bits
__global__ void __launch_bounds__(1024) k_peak_imad(unsigned *out, unsigned m)
{
    unsigned x[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = threadIdx.x * 2654435761u + k;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) {
            unsigned long long p = (unsigned long long)x[k] * m;
            x[k] = (unsigned)(p >> 32) ^ (unsigned)p;
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s ^= x[k];
    if (s == 0x12345u) out[0] = s;
}

// issue rate: four FFMA chains and four integer multiply-add chains (x <- x * x + c: nothing ptxas can
// fold), interleaved -- two pipes, eight independent chains per thread, 64 warps per SM: what limits the
// loop is the four warp schedulers' one instruction per clock each
__global__ void __launch_bounds__(1024) k_peak_issue(unsigned *out, float a, float b, unsigned c)
{
    float f[kChains / 2];
    unsigned x[kChains / 2];
#pragma unroll
    for (int k = 0; k < kChains / 2; ++k) { f[k] = threadIdx.x * 1e-3f + k; x[k] = threadIdx.x * 2654435761u + k; }
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains / 2; ++k) {
            f[k] = fmaf(f[k], a, b);
            x[k] = x[k] * x[k] + c;
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kChains / 2; ++k) s ^= x[k] ^ __float_as_uint(f[k]);
    if (s == 0x12345u) out[0] = s;
}

template <class F>
int time_kernel(F launch, double *ms_out, cudaStream_t st, std::string &err)
{
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { err = "cudaEventCreate failed"; return 1; }
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {          // first two repetitions are warm-up
        cudaEventRecord(a, st);
        launch();
        cudaEventRecord(b, st);
        if (cudaEventSynchronize(b) != cudaSuccess) { err = "peak probe kernel failed"; return 1; }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep >= 2) best = std::min(best, ms);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *ms_out = best;
    count_launch(6);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace

// out[0..6): FP64 FMA TFLOP/s, FP32 FMA TFLOP/s, MUFU Gop/s, DMMA (m8n8k4) TFLOP/s, IMAD.WIDE Gop/s,
// issued warp instructions G/s.  Best of four timed launches each (burst figures: a kernel timed alone).
int probe_peaks(double *out6, cudaStream_t st, std::string &err)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 2, block = 1024;
    void *sink = nullptr;
    if (cudaMalloc(&sink, 64) != cudaSuccess) { err = "cudaMalloc failed"; return 1; }
    const double threads = (double)grid * block, ops = threads * kIters * kChains;
    double ms = 0;
    int rc = 0;
    rc |= time_kernel([&] { k_peak_dfma<<<grid, block, 0, st>>>((double *)sink, 1.0000001, 1e-9); }, &ms, st, err);
    out6[0] = 2.0 * ops / (ms * 1e-3) / 1e12;
    rc |= time_kernel([&] { k_peak_ffma<<<grid, block, 0, st>>>((float *)sink, 1.0000001f, 1e-9f); }, &ms, st, err);
    out6[1] = 2.0 * ops / (ms * 1e-3) / 1e12;
    rc |= time_kernel([&] { k_peak_mufu<<<grid, block, 0, st>>>((float *)sink); }, &ms, st, err);
    out6[2] = ops / (ms * 1e-3) / 1e9;
    rc |= time_kernel([&] { k_peak_dmma<<<grid, block, 0, st>>>((double *)sink, 1.0000001, 1e-9); }, &ms, st, err);
    out6[3] = (threads / 32) * kIters * kChains * 512.0 / (ms * 1e-3) / 1e12;       // 2 * 8 * 8 * 4 flop per warp MMA
    rc |= time_kernel([&] { k_peak_imad<<<grid, block, 0, st>>>((unsigned *)sink, 0xD2511F53u); }, &ms, st, err);
    out6[4] = ops / (ms * 1e-3) / 1e9;
    rc |= time_kernel([&] { k_peak_issue<<<grid, block, 0, st>>>((unsigned *)sink, 1.0000001f, 1e-9f, 12345u); }, &ms, st, err);
    out6[5] = (threads / 32) * kIters * kChains / (ms * 1e-3) / 1e9;
    cudaFree(sink);
    return rc;
}

// DMMA rate against resident warps: out[k] = TFLOP/s with 1 CTA per SM of 128 << k threads (k = 0..3: one, two,
// four, eight warps per scheduler), eight independent accumulator pairs per warp.  Design aid for the Gram
// kernels: how many warps a scheduler needs before the FP64 tensor path stays busy.
int probe_dmma_scaling(double *out4, cudaStream_t st, std::string &err)   // out4: 16 doubles
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    void *sink = nullptr;
    if (cudaMalloc(&sink, 64) != cudaSuccess) { err = "cudaMalloc failed"; return 1; }
    int rc = 0;
    for (int k = 0; k < 4; ++k) {
        const int block = 128 << k;
        double ms = 0;
        rc |= time_kernel([&] { k_peak_dmma<<<sms, block, 0, st>>>((double *)sink, 1.0000001, 1e-9); }, &ms, st, err);
        out4[k] = ((double)sms * block / 32) * kIters * kChains * 512.0 / (ms * 1e-3) / 1e12;
        rc |= time_kernel([&] { k_peak_dmma_operands<<<sms, block, 0, st>>>((double *)sink, 1.0000001, 1e-9); }, &ms, st, err);
        out4[4 + k] = ((double)sms * block / 32) * kIters * kChains * 512.0 / (ms * 1e-3) / 1e12;
        rc |= time_kernel([&] { k_peak_dmma_mul<4><<<sms, block, 0, st>>>((double *)sink, 1.0000001, 1e-9); }, &ms, st, err);
        out4[8 + k] = ((double)sms * block / 32) * kIters * kChains * 512.0 / (ms * 1e-3) / 1e12;
        rc |= time_kernel([&] { k_peak_dmma_mul<8><<<sms, block, 0, st>>>((double *)sink, 1.0000001, 1e-9); }, &ms, st, err);
        out4[12 + k] = ((double)sms * block / 32) * kIters * kChains * 512.0 / (ms * 1e-3) / 1e12;
    }
    cudaFree(sink);
    return rc;
}

}  // namespace bl
