// Shared device helpers for the tuch_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#ifndef TUCH_EXPORT
#define TUCH_EXPORT extern "C" __attribute__((visibility("default")))
#endif

namespace tuch {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define TUCH_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return ::tuch::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define TUCH_REQUIRE(cond, ...)                                                \
    do {                                                                       \
        if (!(cond)) { ::tuch::set_error(__VA_ARGS__); return 1; }             \
    } while (0)

#define TUCH_LAUNCH_CHECK() TUCH_CUDA(cudaGetLastError())

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- MUFU approximations
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// atan2 for the solid-angle sum.  r = min/max in [0,1]; atan(r) = r + r^3 Q(r^2) with a
// degree-6 minimax Q (max abs error 1.1e-7 in fp32, exact linear term so that the ~13k tiny
// far-field angles of a winding number carry no systematic bias).  (+-0, +-0) -> +-0: a query
// that coincides with a triangle corner must contribute nothing (tuch/utils/contact.py:106 sees
// atan2(+-0, +0) there); the IEEE special case atan2(0, -0) = pi is deliberately NOT reproduced.
__device__ __forceinline__ float atan2_poly(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), FLT_MIN);
    const float mn = fminf(ax, ay);
    const float r = mn * rcp_approx(mx);
    const float s = r * r;
    float q = -4.3554012513e-03f;
    q = fmaf(q, s, 2.3040120603e-02f);
    q = fmaf(q, s, -5.7773569451e-02f);
    q = fmaf(q, s, 9.7942332266e-02f);
    q = fmaf(q, s, -1.3976581675e-01f);
    q = fmaf(q, s, 1.9962703910e-01f);
    q = fmaf(q, s, -3.3331659029e-01f);
    float a = fmaf(r * s, q, r);
    a = (ay > ax) ? (1.57079632679489662f - a) : a;
    a = (x < 0.0f) ? (3.14159265358979324f - a) : a;
    return copysignf(a, y);
}

// One (query, triangle) term of the generalized winding number, WITHOUT the factor 2:
// atan2(a.(b x c), |a||b||c| + (a.b)|c| + (a.c)|b| + (b.c)|a|)   (Van Oosterom & Strackee;
// the reference evaluates it tensor-wise in tuch/utils/contact.py:79-106).
__device__ __forceinline__ float half_solid_angle(float px, float py, float pz,
                                                  const float4& A, const float4& B, const float4& C) {
    const float ax = A.x - px, ay = A.y - py, az = A.z - pz;
    const float bx = B.x - px, by = B.y - py, bz = B.z - pz;
    const float cx = C.x - px, cy = C.y - py, cz = C.z - pz;
    const float la = sqrt_approx(fmaf(az, az, fmaf(ay, ay, ax * ax)));
    const float lb = sqrt_approx(fmaf(bz, bz, fmaf(by, by, bx * bx)));
    const float lc = sqrt_approx(fmaf(cz, cz, fmaf(cy, cy, cx * cx)));
    const float crx = fmaf(by, cz, -bz * cy);
    const float cry = fmaf(bz, cx, -bx * cz);
    const float crz = fmaf(bx, cy, -by * cx);
    const float num = fmaf(az, crz, fmaf(ay, cry, ax * crx));
    const float dab = fmaf(az, bz, fmaf(ay, by, ax * bx));
    const float dac = fmaf(az, cz, fmaf(ay, cy, ax * cx));
    const float dbc = fmaf(bz, cz, fmaf(by, cy, bx * cx));
    float den = (la * lb) * lc;          // +0 when the query sits on a corner; keeps den = +0
    den = fmaf(dab, lc, den);
    den = fmaf(dac, lb, den);
    den = fmaf(dbc, la, den);
    return atan2_poly(num, den);
}

// Same term with the numerator taken from the face normal N = (B - A) x (C - A) carried in the w
// components: a.(b x c) is affine in the query and equals N.(A - q), which saves the cross product.
// N.(A - q) is only mathematically 0 when q is another corner of the face; the reference's triple product
// is exactly 0 there and so is the denominator, which is what the guard keys on.  An all-zero padding
// slot contributes exactly 0.
__device__ __forceinline__ float half_solid_angle_n(float px, float py, float pz,
                                                    const float4& A, const float4& B, const float4& C) {
    const float ax = A.x - px, ay = A.y - py, az = A.z - pz;
    const float bx = B.x - px, by = B.y - py, bz = B.z - pz;
    const float cx = C.x - px, cy = C.y - py, cz = C.z - pz;
    const float la = sqrt_approx(fmaf(az, az, fmaf(ay, ay, ax * ax)));
    const float lb = sqrt_approx(fmaf(bz, bz, fmaf(by, by, bx * bx)));
    const float lc = sqrt_approx(fmaf(cz, cz, fmaf(cy, cy, cx * cx)));
    const float num = fmaf(C.w, az, fmaf(B.w, ay, A.w * ax));
    const float dab = fmaf(az, bz, fmaf(ay, by, ax * bx));
    const float dac = fmaf(az, cz, fmaf(ay, cy, ax * cx));
    const float dbc = fmaf(bz, cz, fmaf(by, cy, bx * cx));
    float den = (la * lb) * lc;
    den = fmaf(dab, lc, den);
    den = fmaf(dac, lb, den);
    den = fmaf(dbc, la, den);
    return atan2_poly(den == 0.f ? 0.f : num, den);
}

// ---------------------------------------------------------------- order-independent accumulation
// Gradient scatters add many terms into the same element from different threads.  fp32 atomics make the
// sum depend on the arrival order (run-to-run differences at the 1e-7 level that a discontinuous objective
// amplifies); 64-bit fixed-point atomics are exact and therefore deterministic.
// Representable range: every accumulated SUM must stay within +-2^31 (2.1e9) and terms below 2^-33 (1.2e-10) round
// to zero; a single term beyond +-2^31 saturates (__float2ll_rn) instead of wrapping.  The callers multiply the
// upstream factor (loss weight x g_loss) in BEFORE the quantisation, so "term" means the final gradient
// contribution: with the reference's weights (<= 2000 x tanh'() <= 25/m) sums stay below 1e7.  A non-finite term
// is stored as is into the fp32 destination (it has to poison the result either way) with an atomic exchange.
constexpr float FIX_SCALE = 4294967296.f;            // 2^32
__device__ __forceinline__ void fix_add(long long* acc, float* fallback, float v) {
    if (!isfinite(v)) { atomicExch(fallback, v); return; }
    atomicAdd(reinterpret_cast<unsigned long long*>(acc), (unsigned long long)__float2ll_rn(v * FIX_SCALE));
}
__device__ __forceinline__ float fix_value(long long acc) { return (float)((double)acc * (1.0 / 4294967296.0)); }

// ---------------------------------------------------------------- mbarrier + 1-D TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy through the TMA unit (SASS: UBLKCP); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace tuch
