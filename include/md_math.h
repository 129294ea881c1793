/* md_math.h — the elementary functions of the step path, defined ONCE so that every implementation computes the same
 * bits: sin, cos, atan2 and exp are the only operations on the path that IEEE-754 does not pin (+, -, *, /, sqrt, fmod,
 * floor are exact-rounded or exact in both gcc and nvcc; the CUDA build uses -fmad=false, the C oracle
 * -ffp-contract=off).  With libm's versions the CUDA path and the CPU oracle differ in the last bit of a sine, which a
 * sustained contact amplifies ~100x over a hundred steps; with these the two are bit-identical in every float.
 *
 * Plain float32 polynomial kernels (Cephes single-precision coefficients), Cody-Waite reduction in three parts; accuracy
 * <= 2 ulp on the ranges the path uses (|x| < 1e4 for sin / cos, |x| < 80 for exp), checked against double-precision libm
 * in tests/test_md_math.py.  The reference computes these in Python float64 (math.sin ...): a 1e-7 relative difference,
 * far inside every tolerance of the parity tests. */
#ifndef MD_MATH_H
#define MD_MATH_H
#include <math.h>

#ifdef __CUDACC__
#define MDM_FN __host__ __device__ __forceinline__
#else
#define MDM_FN static inline
#endif

/* Fused multiply-adds are written out explicitly (fmaf is correctly rounded on both sides: FFMA on the GPU, vfmadd or
 * libm on the host); the compilers themselves never contract (-fmad=false / -ffp-contract=off), so both implementations
 * fuse exactly the same products.  These are the building blocks of every dot / cross / matrix product on the path. */
MDM_FN float md_dot3(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(az, bz, fmaf(ay, by, ax * bx)); }
MDM_FN float md_dot4(float aw, float ax, float ay, float az, float bw, float bx, float by, float bz) {
    return fmaf(az, bz, fmaf(ay, by, fmaf(ax, bx, aw * bw)));
}
MDM_FN float md_sum2(float a, float b, float c, float d) { return fmaf(a, b, c * d); }      /* a*b + c*d */
MDM_FN float md_diff2(float a, float b, float c, float d) { return fmaf(a, b, -(c * d)); }  /* a*b - c*d */

#define MDM_PIO2_1 1.5703125f                 /* pi/2 split in three: 8 + 11 + 24 significant bits */
#define MDM_PIO2_2 4.837512969970703125e-4f
#define MDM_PIO2_3 7.54978995489188216e-8f
#define MDM_PI 3.14159265358979323846f

/* sin and cos of the reduced argument r in [-pi/4, pi/4] */
MDM_FN float mdm_sin_k(float r) {
    const float z = r * r;
    return r + r * z * ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f);
}
MDM_FN float mdm_cos_k(float r) {
    const float z = r * r;
    return 1.0f - 0.5f * z + z * z * ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f);
}
MDM_FN float mdm_reduce(float x, int* q) {
    const float kf = floorf(x * 0.63661977236758134308f + 0.5f);   /* nearest multiple of pi/2 */
    *q = (int)kf & 3;
    return ((x - kf * MDM_PIO2_1) - kf * MDM_PIO2_2) - kf * MDM_PIO2_3;
}
MDM_FN float md_sinf(float x) {
    int q;
    const float r = mdm_reduce(x, &q);
    const float v = (q & 1) ? mdm_cos_k(r) : mdm_sin_k(r);
    return (q & 2) ? -v : v;
}
MDM_FN float md_cosf(float x) {
    int q;
    const float r = mdm_reduce(x, &q);
    const float v = (q & 1) ? mdm_sin_k(r) : mdm_cos_k(r);
    return ((q + 1) & 2) ? -v : v;
}
/* atan on [0, inf) */
MDM_FN float mdm_atan_pos(float t) {
    float y0;
    if (t > 2.414213562373095f) { y0 = 1.5707963267948966f; t = -1.0f / t; }
    else if (t > 0.4142135623730950f) { y0 = 0.7853981633974483f; t = (t - 1.0f) / (t + 1.0f); }
    else y0 = 0.0f;
    const float z = t * t;
    return y0 + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * t + t);
}
MDM_FN float md_atan2f(float y, float x) {
    if (x == 0.0f) return y > 0.0f ? 1.5707963267948966f : (y < 0.0f ? -1.5707963267948966f : 0.0f);
    const float t = y / x;
    const float a = t < 0.0f ? -mdm_atan_pos(-t) : mdm_atan_pos(t);
    if (x > 0.0f) return a;
    return y >= 0.0f ? a + MDM_PI : a - MDM_PI;
}
MDM_FN float md_expf(float x) {
    const float kf = floorf(1.44269504088896341f * x + 0.5f);
    float r = x - kf * 0.693359375f;
    r = r - kf * -2.12194440e-4f;
    const float z = r * r;
    const float p = (((((1.9875691500e-4f * r + 1.3981999507e-3f) * r + 8.3334519073e-3f) * r + 4.1665795894e-2f) * r +
                      1.6666665459e-1f) * r + 5.0000001201e-1f) * z + r + 1.0f;
    return ldexpf(p, (int)kf);
}
#endif
