// Counter-based variate sources for the Polya-Gamma engine (sm_100a).
//
// STREAM CONTRACT (DESIGN.md "Stream contract").  Observation i of call c under
// seed s owns the Philox4x32-10 stream with key (s_lo, s_hi) and counter
// (i_lo, i_hi, k, c), k = 0,1,2,..; each block yields four 32-bit words consumed
// in order.  Primitive variates:
//   U : one word w               -> (w + 1/2) 2^-32
//   E : one word (tail-extended) -> -log U, a zero word adds 32 log 2 and redraws
//   N : three words              -> sqrt(-2 log u1) cos(2 pi u2), u1 53-bit, u2 32-bit
//   G : Marsaglia-Tsang from N,U (boost U^(1/a) for a < 1)
// Because the stream is keyed by the GLOBAL observation index, results do not
// depend on the launch geometry, on lane compaction, or on how observations are
// sharded across GPUs.  This replaces the reference's single process-global
// generator (RNG r; LogitWrapper.cpp:68,131).
#pragma once

#include <cstdint>

namespace bl {

struct Philox4x32 {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

    __device__ __forceinline__ static uint4 block(uint4 c, uint2 k)
    {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
            uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += W0;
            k.y += W1;
        }
        return c;
    }
};

__device__ __forceinline__ double word_to_unif(uint32_t w)
{
    return ((double)w + 0.5) * 0x1p-32;
}

// Per-observation Philox stream.
struct PhiloxSource {
    uint2 key;
    uint32_t c0, c1, c3, blk;
    uint4 buf;
    int pos;

    __device__ __forceinline__ void open(uint64_t seed, uint64_t obs, uint32_t call_id)
    {
        key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        c0 = (uint32_t)obs;
        c1 = (uint32_t)(obs >> 32);
        c3 = call_id;
        blk = 0;
        pos = 4;
    }

    // The ten Philox rounds (~70 instructions) stay out of line -- word() has many call sites and
    // the samplers live or die by their code footprint -- but as a PURE function of register
    // arguments (philox_block_ool below): a member taking `this` out of line would pin the whole
    // stream state in local memory, and with ~1000 resident threads those stack frames do not fit
    // the L1 next to the kernels' shared memory (the saddle-point loop spent 60 % of its
    // long-scoreboard stalls on local loads of pos / buf / the counter).
    __device__ __forceinline__ void refill();

    __device__ __forceinline__ uint32_t word()
    {
        if (pos == 4) refill();
        uint32_t w = pos == 0 ? buf.x : pos == 1 ? buf.y : pos == 2 ? buf.z : buf.w;
        ++pos;
        return w;
    }

    // Lazy variates: the handle pins the stream words now; the fp32 estimate (for the
    // decision pre-filters of pg_devroye_fast.cuh) and the fp64 value can be had later.
    struct LazyE { uint32_t w; int k; };
    struct LazyN { uint32_t w0, w1, w2; };

    __device__ __forceinline__ LazyE expon_lazy()
    {
        LazyE e{word(), 0};
        while (e.w == 0u) { ++e.k; e.w = word(); }
        return e;
    }
    __device__ __forceinline__ static float approx(const LazyE &e)
    {
        return (float)e.k * 22.18070977791825f - __logf(((float)e.w + 0.5f) * 0x1p-32f);
    }
    __device__ __forceinline__ static double exact(const LazyE &e)
    {
        return (double)e.k * (32.0 * 0.693147180559945309417232) - log(word_to_unif(e.w));
    }
    __device__ __forceinline__ LazyN norm_lazy() { return LazyN{word(), word(), word()}; }
    __device__ __forceinline__ static float approx(const LazyN &n)
    {
        float u1 = ((float)n.w0 + 0.5f) * 0x1p-32f;   // top 32 of the 53 radius bits
        float u2 = ((float)n.w2 + 0.5f) * 0x1p-32f;
        return sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2);
    }
    __device__ __forceinline__ static double exact(const LazyN &n)
    {
        uint64_t m = ((uint64_t)n.w0 << 21) | (uint64_t)(n.w1 >> 11);
        double u1 = ((double)m + 0.5) * 0x1p-53;
        double u2 = word_to_unif(n.w2);
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }

    __device__ __forceinline__ double unif() { return word_to_unif(word()); }

    // the transcendental part of E and N out of line (pure functions of the pinned words)
    __device__ __forceinline__ double expon();
    __device__ __forceinline__ double norm();

    __device__ double gamma(double a)
    {
        double boost = 1.0;
        if (a < 1.0) {
            boost = exp(log(unif()) / a);
            a += 1.0;
        }
        double d = a - 1.0 / 3.0;
        double c = 1.0 / sqrt(9.0 * d);
        for (;;) {
            double x = norm();
            double u = unif();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
        }
    }

    __device__ __forceinline__ bool exhausted() const { return false; }
    __device__ __forceinline__ void counts(int *t) const { t[0] = t[1] = t[2] = t[3] = 0; }
};

static __device__ __noinline__ uint4 philox_block_ool(uint32_t c0, uint32_t c1, uint32_t blk, uint32_t c3, uint2 key)
{
    return Philox4x32::block(make_uint4(c0, c1, blk, c3), key);
}
static __device__ __noinline__ double philox_expon_ool(uint32_t w, int k)
{
    return PhiloxSource::exact(PhiloxSource::LazyE{w, k});
}
static __device__ __noinline__ double philox_norm_ool(uint32_t w0, uint32_t w1, uint32_t w2)
{
    return PhiloxSource::exact(PhiloxSource::LazyN{w0, w1, w2});
}

__device__ __forceinline__ void PhiloxSource::refill()
{
    buf = philox_block_ool(c0, c1, blk, c3, key);
    ++blk;
    pos = 0;
}
__device__ __forceinline__ double PhiloxSource::expon()
{
    LazyE e = expon_lazy();
    return philox_expon_ool(e.w, e.k);
}
__device__ __forceinline__ double PhiloxSource::norm()
{
    LazyN n = norm_lazy();
    return philox_norm_ool(n.w0, n.w1, n.w2);
}

// Injected variate tape (tier-1 parity): one segment per observation and kind,
// consumed in the reference's statement order.  A dry segment flags the draw as
// exhausted and keeps rejection loops finite by falling back to a fixed stream.
struct TapeSource {
    const double *tu, *te, *tn, *tg;
    int lu, le, ln, lg;
    int cu, ce, cn, cg;
    bool dry;
    PhiloxSource fb;

    __device__ void open(const double *tu_, int lu_, const double *te_, int le_,
                         const double *tn_, int ln_, const double *tg_, int lg_, size_t i)
    {
        tu = tu_ ? tu_ + i * (size_t)lu_ : nullptr; lu = tu_ ? lu_ : 0;
        te = te_ ? te_ + i * (size_t)le_ : nullptr; le = te_ ? le_ : 0;
        tn = tn_ ? tn_ + i * (size_t)ln_ : nullptr; ln = tn_ ? ln_ : 0;
        tg = tg_ ? tg_ + i * (size_t)lg_ : nullptr; lg = tg_ ? lg_ : 0;
        cu = ce = cn = cg = 0;
        dry = false;
        fb.open(0x243F6A889E3779B9ull, 0, 0);
    }

    __device__ double unif()
    {
        int k = cu++;
        if (k >= lu) { dry = true; return fb.unif(); }
        return tu[k];
    }
    __device__ double expon()
    {
        int k = ce++;
        if (k >= le) { dry = true; return fb.expon(); }
        return te[k];
    }
    __device__ double norm()
    {
        int k = cn++;
        if (k >= ln) { dry = true; return 2.0 * fb.unif() - 1.0; }
        return tn[k];
    }
    __device__ double gamma(double)
    {
        int k = cg++;
        if (k >= lg) { dry = true; return fb.expon(); }
        return tg[k];
    }
    struct LazyE { double v; };
    struct LazyN { double v; };
    __device__ LazyE expon_lazy() { return LazyE{expon()}; }
    __device__ LazyN norm_lazy() { return LazyN{norm()}; }
    __device__ static float approx(const LazyE &e) { return (float)e.v; }
    __device__ static double exact(const LazyE &e) { return e.v; }
    __device__ static float approx(const LazyN &n) { return (float)n.v; }
    __device__ static double exact(const LazyN &n) { return n.v; }

    __device__ bool exhausted() const { return dry; }
    __device__ void counts(int *t) const { t[0] = cu; t[1] = ce; t[2] = cn; t[3] = cg; }
};

}  // namespace bl
