/* TEST INFRASTRUCTURE ONLY -- included twice by adrt_oracle.c with
 *   T   = float / double
 *   FN(name) = name##_f32 / name##_f64
 *
 * Plain-C restatement of the reference's numerical core in the public
 * "Q-layout" (B,4,D,n) with D = 2n-1 (index [b][q][d][c], c contiguous).
 * Every function cites the reference file:line (relative to
 * /root/reference/src/adrt/) whose arithmetic it follows.  All sums are the
 * same two-operand IEEE adds in the same order; compiled with
 * -ffp-contract=off so nothing is fused.
 */

/* core.py:169-176 (adrt_init) == adrt_cdefs_adrt.hpp:124-186 (load phase of
 * adrt_basic): four oriented copies of the image in rows 0..n-1, zeros below. */
void FN(oracle_adrt_init)(const T *x, long B, long n, T *out)
{
    const long D = 2 * n - 1;
    for (long b = 0; b < B; ++b) {
        const T *img = x + b * n * n;
        T *o = out + b * 4 * D * n;
        for (long q = 0; q < 4; ++q)
            for (long d = 0; d < D; ++d)
                for (long c = 0; c < n; ++c) {
                    T v = 0;
                    if (d < n) {
                        switch (q) {
                        case 0: v = img[c * n + (n - 1 - d)]; break;           /* flip cols, transpose */
                        case 1: v = img[(n - 1 - d) * n + c]; break;           /* flip rows            */
                        case 2: v = img[d * n + c]; break;                     /* identity             */
                        default: v = img[(n - 1 - c) * n + (n - 1 - d)]; break; /* flip both, transpose */
                        }
                    }
                    o[(q * D + d) * n + c] = v;
                }
    }
}

/* adrt_cdefs_adrt.hpp:215-258 (adrt_step); same arithmetic as adrt_core
 * (adrt_cdefs_adrt.hpp:55-96) transposed into the Q-layout.  One add per
 * output, or a plain copy where the shifted operand does not exist. */
void FN(oracle_adrt_step)(const T *in, long B, long n, int iter, T *out)
{
    const long D = 2 * n - 1;
    const long e = 1L << iter, e2 = 2 * e;
    for (long p = 0; p < B * 4; ++p) {
        const T *I = in + p * D * n;
        T *O = out + p * D * n;
        for (long d = 0; d < D; ++d)
            for (long c = 0; c < n; ++c) {
                const long k = c / e2, a = c % e2;
                const long cA = 2 * k * e + a / 2;
                const long cB = (2 * k + 1) * e + a / 2;
                const long sh = (a + 1) / 2;
                if (d >= sh) O[d * n + c] = I[d * n + cA] + I[(d - sh) * n + cB];
                else         O[d * n + c] = I[d * n + cA];
            }
    }
}

/* adrt_cdefs_bdrt.hpp:190-244 (bdrt_step): missing operands are literal +0
 * that still take part in the add (aval = 0; bval = 0; aval + bval). */
void FN(oracle_bdrt_step)(const T *in, long B, long n, int iter, int K, T *out)
{
    const long D = 2 * n - 1;
    const long e = 1L << (K - 1 - iter);
    for (long p = 0; p < B * 4; ++p) {
        const T *I = in + p * D * n;
        T *O = out + p * D * n;
        for (long d = 0; d < D; ++d)
            for (long c = 0; c < n; ++c) {
                const long cb = c / e, ci = c % e;
                const long beta = 2 * (ci + e * (cb / 2));
                if (cb % 2 == 0) {
                    O[d * n + c] = I[d * n + beta] + I[d * n + beta + 1];
                } else {
                    const long r = d + ci;
                    T av = 0, bv = 0;
                    if (r < D) av = I[r * n + beta];
                    if (r + 1 < D) bv = I[(r + 1) * n + beta + 1];
                    O[d * n + c] = av + bv;
                }
            }
    }
}

/* adrt_cdefs_bdrt.hpp:55-116 (bdrt_core as used by bdrt_basic), in Q-layout:
 * identical to bdrt_step except that the last valid row of a shifted section
 * is a COPY of la_val (bdrt.hpp:96-103) and the rows after it are literal
 * zeros (bdrt.hpp:105-109) -- observable only on negative zeros. */
static void FN(oracle_bdrt_core_q)(const T *in, long B, long n, int iter, int K, T *out)
{
    const long D = 2 * n - 1;
    const long e = 1L << (K - 1 - iter);
    for (long p = 0; p < B * 4; ++p) {
        const T *I = in + p * D * n;
        T *O = out + p * D * n;
        for (long d = 0; d < D; ++d)
            for (long c = 0; c < n; ++c) {
                const long cb = c / e, ci = c % e;
                const long beta = 2 * (ci + e * (cb / 2));
                if (cb % 2 == 0) {
                    O[d * n + c] = I[d * n + beta] + I[d * n + beta + 1];
                } else {
                    const long r = d + ci;
                    if (r + 1 < D)      O[d * n + c] = I[r * n + beta] + I[(r + 1) * n + beta + 1];
                    else if (r < D)     O[d * n + c] = I[r * n + beta];
                    else                O[d * n + c] = 0;
                }
            }
    }
}

/* adrt_cdefs_adrt.hpp:101-212 (adrt_basic) = init + K stages.  tmp must hold
 * B*4*D*n elements. */
void FN(oracle_adrt)(const T *x, long B, long n, int K, T *tmp, T *out)
{
    T *a = (K % 2 == 0) ? out : tmp, *b = (K % 2 == 0) ? tmp : out;
    FN(oracle_adrt_init)(x, B, n, a);
    for (int i = 0; i < K; ++i) {
        FN(oracle_adrt_step)(a, B, n, i, b);
        T *t = a; a = b; b = t;
    }
    /* result is in `a`, which is `out` by construction */
}

/* adrt_cdefs_bdrt.hpp:121-187 (bdrt_basic) = K transposed stages. */
void FN(oracle_bdrt)(const T *y, long B, long n, int K, T *tmp, T *out)
{
    const long N = B * 4 * (2 * n - 1) * n;
    if (K == 0) { for (long i = 0; i < N; ++i) out[i] = y[i]; return; }
    const T *src = y;
    T *a = (K % 2 == 1) ? out : tmp, *b = (K % 2 == 1) ? tmp : out;
    for (int i = 0; i < K; ++i) {
        FN(oracle_bdrt_core_q)(src, B, n, i, K, a);
        src = a;
        T *t = a; a = b; b = t;
    }
}

/* adrt_cdefs_iadrt.hpp:52-105 (iadrt_core) + :110-176 (iadrt_basic), in the
 * Q-layout: column index of buffer entry (l, col) at a stage with C columns
 * per l is l*C + col.  Evaluation order ((0 + A) - B) + prev as in :74-97. */
void FN(oracle_iadrt)(const T *y, long B, long n, int K, T *tmp, T *out)
{
    const long D = 2 * n - 1, N = B * 4 * D * n;
    if (K == 0) { for (long i = 0; i < N; ++i) out[i] = y[i]; return; }
    const T *src = y;
    T *a = (K % 2 == 1) ? out : tmp, *b = (K % 2 == 1) ? tmp : out;
    for (int s = 0; s < K; ++s) {
        const long Cin = n >> s, C = Cin / 2, L = 2L << s;
        for (long p = 0; p < B * 4; ++p) {
            const T *I = src + p * D * n;
            T *O = a + p * D * n;
            for (long l = 0; l < L; ++l)
                for (long col = 0; col < C; ++col) {
                    const long A = (l / 2) * Cin + 2 * col, co = l * C + col;
                    for (long d = D - 1; d >= 0; --d) {
                        T val = 0;
                        if (l % 2 == 0) {
                            val += I[d * n + A];
                            if (d + 1 < D) val -= I[(d + 1) * n + A + 1];
                        } else if (d + 1 + col < D) {
                            val += I[(d + 1 + col) * n + A + 1];
                            val -= I[(d + 1 + col) * n + A];
                        }
                        if (d + 1 < D) val += O[(d + 1) * n + co];
                        O[d * n + co] = val;
                    }
                }
        }
        src = a;
        T *t = a; a = b; b = t;
    }
}

/* adrt_cdefs_fmg.hpp:53-73 */
void FN(oracle_fmg_restriction)(const T *in, long B, long n, T *out)
{
    const long D = 2 * n - 1, R = n - 1, C = n / 2;
    for (long p = 0; p < B * 4; ++p)
        for (long r = 0; r < R; ++r)
            for (long c = 0; c < C; ++c) {
                const T va = in[(p * D + 2 * r) * n + 2 * c];
                const T vb = in[(p * D + 2 * r + 1) * n + 2 * c];
                out[(p * R + r) * C + c] = (va + vb) / (T)4;
            }
}

/* adrt_cdefs_fmg.hpp:75-95 */
void FN(oracle_fmg_prolongation)(const T *in, long B, long h, long w, T *out)
{
    for (long b = 0; b < B; ++b)
        for (long r = 0; r < h; ++r)
            for (long c = 0; c < w; ++c) {
                const T v = in[(b * h + r) * w + c];
                T *o = out + b * 4 * h * w;
                o[(2 * r) * 2 * w + 2 * c] = v;
                o[(2 * r) * 2 * w + 2 * c + 1] = v;
                o[(2 * r + 1) * 2 * w + 2 * c] = v;
                o[(2 * r + 1) * 2 * w + 2 * c + 1] = v;
            }
}

/* adrt_cdefs_fmg.hpp:97-171: reflect-101 boundary, each product rounded
 * separately, sum order (v11+v21+v31)+(v12+v22+v32)+(v13+v23+v33). */
void FN(oracle_fmg_highpass)(const T *in, long B, long h, long w, T *out)
{
    const T ca = (T)-0.0625L, cb = (T)-0.125L, cc = (T)0.75L;
    for (long b = 0; b < B; ++b) {
        const T *I = in + b * h * w;
        T *O = out + b * h * w;
        for (long r = 0; r < h; ++r) {
            const long pr = (r == 0 ? 1 : r - 1), nr = (r == h - 1 ? r - 1 : r + 1);
            for (long c = 0; c < w; ++c) {
                const long pc = (c == 0 ? 1 : c - 1), nc = (c == w - 1 ? c - 1 : c + 1);
                const T v11 = ca * I[pr * w + pc], v12 = cb * I[pr * w + c], v13 = ca * I[pr * w + nc];
                const T v21 = cb * I[r * w + pc],  v22 = cc * I[r * w + c],  v23 = cb * I[r * w + nc];
                const T v31 = ca * I[nr * w + pc], v32 = cb * I[nr * w + c], v33 = ca * I[nr * w + nc];
                O[r * w + c] = (v11 + v21 + v31) + (v12 + v22 + v32) + (v13 + v23 + v33);
            }
        }
    }
}

/* adrt_cdefs_interp_adrtcart.hpp:61-114 with float_index = float
 * (adrt_cdefs_py.cpp:638).  lerp follows libstdc++'s std::lerp for the
 * opposite-sign case (t*b + (1-t)*a); tanf/cosf/roundf/floorf/sqrt are the
 * platform libm, exactly what the reference links. */
void FN(oracle_interp_to_cart)(const T *in, long B, long n, T *out)
{
    const long D = 2 * n - 1, W = 4 * n;
    const float sqrt2_2 = (float)1.41421356237309504880168872420969808L / 2.0f;
    const float pi = (float)3.14159265358979323846264338327950288L;
    const float pi_2 = pi / 2.0f, pi_4 = pi / 4.0f, pi_8 = pi / 8.0f;
    const float t_left = sqrt2_2 - (sqrt2_2 / (float)n);
    const float th_left = pi_2 - (pi_8 / (float)n);
    for (long b = 0; b < B; ++b)
        for (long off = 0; off < n; ++off)
            for (long ang = 0; ang < W; ++ang) {
                const float of = (float)off / (float)(n - 1);
                const float af = (float)ang / (float)(W - 1);
                const float t = t_left * (of * 1.0f + (1.0f - of) * -1.0f);
                const float th = th_left * (af * -1.0f + (1.0f - af) * 1.0f);
                float qf = -th / pi_4;
                qf = qf < -2.0f ? -2.0f : (qf > 1.0f ? 1.0f : qf);
                const int q = (int)(floorf(qf) + 2);
                const int sgn = (q % 2 == 0) ? 1 : -1;
                const float th0 = pi_4 - fabsf(fabsf(th) - pi_4);
                float tant = tanf(th0);
                tant = tant < 0.0f ? 0.0f : (tant > 1.0f ? 1.0f : tant);
                const float si = roundf(tant * (float)(n - 1));
                T factor;
                if (sizeof(T) > sizeof(float)) {
                    const double sa = (double)si / (double)(n - 1);
                    factor = (T)sqrt(sa * sa + 1.0);
                } else {
                    const float sa = si / (float)(n - 1);
                    factor = (T)sqrtf(sa * sa + 1.0f);
                }
                const float h0 = (0.5f + (tant / 2.0f)) + ((sgn >= 0 ? t : -t) / cosf(th0));
                const float hi = (roundf(h0 * (float)(2 * n)) - 1.0f) / 2.0f;
                T v = 0;
                if (hi >= 0.0f && hi < (float)D)
                    v = factor * in[((b * 4 + q) * D + (long)hi) * n + (long)si];
                out[(b * n + off) * W + ang] = v;
            }
}
