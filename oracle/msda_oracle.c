/*
 * Plain-C restatement of the MSDeformAttn core.  TEST INFRASTRUCTURE ONLY — never linked
 * into or loaded by the product library (libcape_msda.so).
 *
 * Follows /root/reference/models/deformable_transformer.py:115-141
 * (ms_deform_attn_core_pytorch): sampling_grids = 2*loc-1 (:129), bilinear grid_sample with
 * zeros padding and align_corners=False per level (:136-137), attention-weighted sum over the
 * L*P samples (:139-140), output (N, Lq, M*D) (:141).  The backward is the analytic derivative
 * (grid_sampler_2d_backward + the product rule for the weights); formulas in oracle/msda_numpy.py.
 *
 * Built by oracle/build_oracle.py:  gcc -O2 -fopenmp -shared -fPIC  ->  oracle/_build/libmsda_oracle.so
 * Threads: OpenMP over (n, q) in the forward and over (n, m) in the backward — every (n, m)
 * pair owns a disjoint slice of grad_value, so no atomics and a deterministic summation order.
 *
 * The body is instantiated twice (REAL = float, accumulating in double; REAL = double).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int msda_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define DEFINE_ORACLE(SUFFIX, REAL)                                                                  \
                                                                                                     \
/* one sample: pixel coordinates, corner rows, weights; returns a 4-bit in-bounds mask */            \
static inline int corners_##SUFFIX(REAL locx, REAL locy, int64_t H, int64_t W, int64_t start,        \
                                   int64_t rows[4], double wx[2], double wy[2]) {                    \
    /* :129 then grid_sample's unnormalise for align_corners=False, in REAL like ATen */             \
    REAL gx = (REAL)2 * locx - (REAL)1, gy = (REAL)2 * locy - (REAL)1;                               \
    REAL x = (gx + (REAL)1) * (REAL)((double)W / 2.0) - (REAL)0.5;                                   \
    REAL y = (gy + (REAL)1) * (REAL)((double)H / 2.0) - (REAL)0.5;                                   \
    if (!(x >= (REAL)-1 && x < (REAL)W && y >= (REAL)-1 && y < (REAL)H)) return 0; /* no corner in bounds */                     \
    REAL x0f = (REAL)floor((double)x), y0f = (REAL)floor((double)y);                                 \
    REAL lx = x - x0f, ly = y - y0f;                                                                 \
    int64_t x0 = (int64_t)x0f, y0 = (int64_t)y0f;                                                    \
    wx[0] = (double)((REAL)1 - lx); wx[1] = (double)lx;                                              \
    wy[0] = (double)((REAL)1 - ly); wy[1] = (double)ly;                                              \
    int mask = 0;                                                                                    \
    for (int c = 0; c < 4; ++c) {                                                                    \
        int64_t xi = x0 + (c & 1), yi = y0 + (c >> 1);                                               \
        if (xi >= 0 && xi < W && yi >= 0 && yi < H) {                                                \
            mask |= 1 << c;                                                                          \
            rows[c] = start + yi * W + xi;                                                           \
        } else {                                                                                     \
            rows[c] = start;                                                                         \
        }                                                                                            \
    }                                                                                                \
    return mask;                                                                                     \
}                                                                                                    \
                                                                                                     \
int msda_oracle_forward_##SUFFIX(const REAL* value, const int64_t* shapes, const int64_t* starts,    \
                                 const REAL* loc, const REAL* attn, REAL* out,                       \
                                 int N, int S, int M, int D, int Lq, int L, int P) {                 \
    if (D > 1024 || L > 64) return -1;                                                               \
    const int64_t nq = (int64_t)N * Lq;                                                              \
    _Pragma("omp parallel for schedule(static)")                                                     \
    for (int64_t t = 0; t < nq; ++t) {                                                               \
        const int64_t n = t / Lq;                                                                    \
        double acc[1024];                                                                            \
        for (int m = 0; m < M; ++m) {                                                                \
            for (int d = 0; d < D; ++d) acc[d] = 0.0;                                                \
            const REAL* lp = loc + ((t * M + m) * (int64_t)L * P) * 2;                               \
            const REAL* ap = attn + (t * M + m) * (int64_t)L * P;                                    \
            for (int l = 0; l < L; ++l) {                                                            \
                const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];                              \
                for (int p = 0; p < P; ++p) {                                                        \
                    int64_t rows[4]; double wx[2], wy[2];                                            \
                    const int mask = corners_##SUFFIX(lp[(l * P + p) * 2], lp[(l * P + p) * 2 + 1],  \
                                                      H, W, starts[l], rows, wx, wy);                \
                    const double a = (double)ap[l * P + p];                                          \
                    for (int c = 0; c < 4; ++c) {                                                    \
                        if (!(mask >> c & 1)) continue;                                              \
                        const double w = a * wx[c & 1] * wy[c >> 1];                                 \
                        const REAL* v = value + ((n * S + rows[c]) * M + m) * (int64_t)D;            \
                        for (int d = 0; d < D; ++d) acc[d] += w * (double)v[d];                      \
                    }                                                                                \
                }                                                                                    \
            }                                                                                        \
            REAL* o = out + (t * M + m) * (int64_t)D;                                                \
            for (int d = 0; d < D; ++d) o[d] = (REAL)acc[d];                                         \
        }                                                                                            \
    }                                                                                                \
    return 0;                                                                                        \
}                                                                                                    \
                                                                                                     \
int msda_oracle_backward_##SUFFIX(const REAL* gout, const REAL* value, const int64_t* shapes,        \
                                  const int64_t* starts, const REAL* loc, const REAL* attn,          \
                                  REAL* gvalue, REAL* gloc, REAL* gattn,                             \
                                  int N, int S, int M, int D, int Lq, int L, int P) {                \
    if (D > 1024 || L > 64) return -1;                                                               \
    const int64_t nm = (int64_t)N * M;                                                               \
    _Pragma("omp parallel for schedule(dynamic, 1)")                                                 \
    for (int64_t t = 0; t < nm; ++t) {                                                               \
        const int64_t n = t / M; const int m = (int)(t % M);                                         \
        /* private double accumulator for this (n, m) slice of grad_value */                         \
        double* gv = (double*)__builtin_malloc((size_t)S * D * sizeof(double));                      \
        memset(gv, 0, (size_t)S * D * sizeof(double));                                               \
        for (int64_t q = 0; q < Lq; ++q) {                                                           \
            const int64_t qm = (n * Lq + q) * M + m;                                                 \
            const REAL* g = gout + qm * (int64_t)D;                                                  \
            const REAL* lp = loc + qm * (int64_t)L * P * 2;                                          \
            const REAL* ap = attn + qm * (int64_t)L * P;                                             \
            for (int l = 0; l < L; ++l) {                                                            \
                const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];                              \
                for (int p = 0; p < P; ++p) {                                                        \
                    int64_t rows[4]; double wx[2], wy[2];                                            \
                    const int mask = corners_##SUFFIX(lp[(l * P + p) * 2], lp[(l * P + p) * 2 + 1],  \
                                                      H, W, starts[l], rows, wx, wy);                \
                    const double a = (double)ap[l * P + p];                                          \
                    double ga = 0.0, gx = 0.0, gy = 0.0;                                             \
                    for (int c = 0; c < 4; ++c) {                                                    \
                        if (!(mask >> c & 1)) continue;                                              \
                        const REAL* v = value + ((n * S + rows[c]) * M + m) * (int64_t)D;            \
                        double* gvr = gv + (rows[c]) * (int64_t)D;                                   \
                        const double w = wx[c & 1] * wy[c >> 1];                                     \
                        double dot = 0.0;                                                            \
                        for (int d = 0; d < D; ++d) {                                                \
                            dot += (double)g[d] * (double)v[d];                                      \
                            gvr[d] += a * w * (double)g[d];                                          \
                        }                                                                            \
                        ga += w * dot;                                                               \
                        gx += ((c & 1) ? wy[c >> 1] : -wy[c >> 1]) * dot;                            \
                        gy += ((c >> 1) ? wx[c & 1] : -wx[c & 1]) * dot;                             \
                    }                                                                                \
                    gattn[qm * (int64_t)L * P + l * P + p] = (REAL)ga;                               \
                    gloc[(qm * (int64_t)L * P + l * P + p) * 2] = (REAL)(a * (double)W * gx);        \
                    gloc[(qm * (int64_t)L * P + l * P + p) * 2 + 1] = (REAL)(a * (double)H * gy);    \
                }                                                                                    \
            }                                                                                        \
        }                                                                                            \
        for (int64_t s = 0; s < S; ++s) {                                                            \
            REAL* dst = gvalue + ((n * S + s) * M + m) * (int64_t)D;                                 \
            for (int d = 0; d < D; ++d) dst[d] = (REAL)gv[s * D + d];                                \
        }                                                                                            \
        __builtin_free(gv);                                                                          \
    }                                                                                                \
    return 0;                                                                                        \
}

DEFINE_ORACLE(f32, float)
DEFINE_ORACLE(f64, double)
