/* validate_fast_arith.c — randomized CPU check of the two arithmetic shortcuts of abea_fill_kernel<FAST>
 * (f5c_b200/csrc/abea_kernels.cuh) against the reference's plain IEEE expressions (reference src/align.c:108-115,
 * 137-152, 378-392), over exactly the input ranges abea_prepare_kernel / abea_load_kernel admit to the FAST path:
 *
 *   (1) emission:  a = (x - m) / s ; lp = lead + ((-0.5f * a) * a)      [reference]
 *                  t = x - m ; q0 = t * r ; rem = fma(-s, q0, t) ; a = fma(rem, r, q0) ; lp = fma(a * a, -0.5f, lead)
 *                  with r = RN(1 / s)                                    [kernel]
 *       x, m in {0} U [2^-60, 2^16] (either sign), s in [2^-6, 2^12] with a mantissa that is not all ones
 *   (2) rounding a double sum to float precision inside the FP64 pipe: (x + C) - C with
 *       C = sign(x) * 1.5 * 2^(e + 29), e = exponent of x, low word of C taken from a float-valued "donor"
 *       vs (double)(float)x, for sums of the shape the DP forms (float-valued + double constant + float-valued)
 *
 * Build and run (tools/, test infrastructure; prints mismatch counts, exit status 1 on any mismatch):
 *   gcc -O2 -ffp-contract=off -o /tmp/validate_fast_arith tools/validate_fast_arith.c -lm && /tmp/validate_fast_arith 400
 *   (argument: millions of trials per test, default 100)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s_[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static inline uint64_t rnd(void) { /* xoroshiro128+ */
    uint64_t a = s_[0], b = s_[1], r = a + b;
    b ^= a;
    s_[0] = ((a << 24) | (a >> 40)) ^ b ^ (b << 16);
    s_[1] = (b << 37) | (b >> 27);
    return r;
}
static inline float f_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t bits_from_f(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline double d_from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t bits_from_d(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

/* random float with exponent uniform in [elo, ehi] (unbiased), random mantissa and sign */
static float rnd_float(int elo, int ehi) {
    uint64_t r = rnd();
    int e = elo + (int)(r % (uint64_t)(ehi - elo + 1));
    uint32_t man = (uint32_t)(r >> 20) & 0x7fffffu;
    uint32_t sign = (uint32_t)(r >> 63);
    return f_from_bits((sign << 31) | ((uint32_t)(e + 127) << 23) | man);
}

static int test_emission(long trials) {
    long bad = 0, zero_t = 0;
    for (long i = 0; i < trials; i++) {
        uint64_t sel = rnd();
        float x, m, s, lead;
        /* level means as the models have them (2^5..2^8), or anything admitted */
        m = (sel & 1) ? rnd_float(5, 8) : rnd_float(-60, 15);
        if ((sel & 0xff00) == 0) m = 0.0f;
        switch ((sel >> 1) & 3) {
        case 0: x = rnd_float(-60, 15); break;                 /* anything admitted */
        case 1: x = rnd_float(-60, -7); break;                 /* the tiny means that used to be rejected */
        case 2: x = m + (float)((int)((sel >> 32) & 0xffff) - 32768) * f_from_bits(bits_from_f(fabsf(m) + 1e-30f) & 0x7f800000u) * 1.1920929e-7f; break; /* within a few ulps of m */
        default: x = m * (1.0f + (float)(int)((sel >> 40) & 0xff) / 64.0f); break;
        }
        if ((sel & 0xff0000) == 0) x = 0.0f;
        {
            float ax = fabsf(x), am = fabsf(m);
            if (!(ax == 0.0f || (ax >= 0x1p-60f && ax <= 65536.0f))) continue;
            if (!(am == 0.0f || (am >= 0x1p-60f && am <= 65536.0f))) continue;
        }
        s = fabsf(rnd_float(-6, 11));
        if ((bits_from_f(s) & 0x7fffffu) == 0x7fffffu) continue;
        lead = -0.918938f - logf(s);
        /* reference */
        volatile float t0 = x - m;
        volatile float a0 = t0 / s;
        volatile float h0 = -0.5f * a0;
        volatile float p0 = h0 * a0;
        float lp0 = lead + p0;
        /* kernel */
        float r = 1.0f / s;
        volatile float t1 = x - m;
        volatile float q0 = t1 * r;
        float rem = fmaf(-s, q0, t1);
        float a1 = fmaf(rem, r, q0);
        volatile float sq = a1 * a1;
        float lp1 = fmaf(sq, -0.5f, lead);
        if (t0 == 0.0f) zero_t++;
        if (bits_from_f(lp0) != bits_from_f(lp1) && !(lp0 == 0.0f && lp1 == 0.0f)) {
            if (bad < 10) fprintf(stderr, "emission mismatch: x=%a m=%a s=%a ref=%a fast=%a (a %a vs %a)\n", x, m, s, lp0, lp1, a0, a1);
            bad++;
        }
    }
    printf("emission: %ld trials, %ld with x == m, %ld mismatches\n", trials, zero_t, bad);
    return bad != 0;
}

static double round_fast(double x, double donor) {
    uint64_t xb = bits_from_d(x);
    uint32_t hi = (uint32_t)(xb >> 32);
    uint32_t chi = (hi >> 20) * 0x00100000u + 0x01d80000u;
    double C = d_from_bits(((uint64_t)chi << 32) | (uint32_t)bits_from_d(donor));
    volatile double y = x + C;
    return y - C;
}

static int test_rounding(long trials) {
    long bad = 0;
    for (long i = 0; i < trials; i++) {
        /* a DP sum: float-valued score + double transition constant + float-valued emission */
        float prev = -fabsf(rnd_float(-3, 14));
        float lp = rnd_float(-8, 9);
        double cst = -fabs((double)rnd_float(-8, 5)) + (double)rnd_float(-40, -30); /* 53-bit-ish constant */
        if ((rnd() & 0xff) == 0) prev = 0.0f;
        volatile double s1 = (double)prev + cst;
        volatile double x = s1 + (double)lp;
        if (rnd() & 1) { /* exact ties at 24 bits: a float-valued number plus half an ulp */
            float base = -fabsf(rnd_float(-3, 14));
            x = (double)base + ldexp(1.0, ilogbf(base) - 24) * ((rnd() & 1) ? 1.0 : -1.0);
        }
        double want = (double)(float)x;
        double got = round_fast(x, (double)prev);
        if (bits_from_d(want) != bits_from_d(got) && !(want == 0.0 && got == 0.0)) {
            if (bad < 10) fprintf(stderr, "rounding mismatch: x=%a want=%a got=%a\n", (double)x, want, got);
            bad++;
        }
    }
    printf("rounding: %ld trials, %ld mismatches\n", trials, bad);
    return bad != 0;
}

/* (3) one DP cell with ONE rounding (abea_cell_once): the reference rounds the three candidate sums to float and compares
 *     the floats, ties L > U > D (src/align.c:378-392). Rounding is monotone, so the cell's score is RN(max of the three
 *     unrounded sums); a candidate ties with (or is) the maximum exactly when it lies in the rounding interval of that
 *     score, i.e. at or above its lower end lb — |R| + half an ulp of float, inclusive when R's last mantissa bit is even
 *     — which for a NEGATIVE float-valued double R is an integer operation on its low word. All cell scores of a FAST
 *     read are negative (every emission constant -0.918938 - log stdv is <= 0: checked per read) or -inf. */
static void cell_ref(double lpd, double up, double left, double diag, double lp_step, double lp_stay, double lp_skip,
                     double* score, int* from) {
    volatile double d0 = diag + lp_step, u0 = up + lp_stay;
    volatile float sd = (float)(d0 + lpd), su = (float)(u0 + lpd), sl = (float)(left + lp_skip);
    float mx = sd;
    int f = 0;
    if (su > mx) mx = su;
    if (mx == su) f = 1;
    if (sl > mx) mx = sl;
    if (mx == sl) f = 2;
    *score = (double)mx;
    *from = f;
}
static void cell_once(double lpd, double up, double left, double diag, double lp_step, double lp_stay, double lp_skip,
                      double* score, int* from) {
    volatile double d0 = diag + lp_step, u0 = up + lp_stay;
    volatile double d = d0 + lpd, u = u0 + lpd, l = left + lp_skip;
    double m2 = (u >= d) ? u : d;
    double m3 = (l >= m2) ? l : m2;
    double R = (double)(float)m3;
    uint64_t rb = bits_from_d(R);
    uint32_t lo = (uint32_t)rb;
    uint32_t lb_lo = lo + 0x10000000u - ((lo >> 29) & 1u);
    double lb = d_from_bits((rb & 0xffffffff00000000ull) | lb_lo);
    int isL = !(l < lb), isU = !(u < lb);
    *score = R;
    *from = isL ? 2 : (isU ? 1 : 0);
}
static double neg_score(void) { /* a float-valued double: a negative band score, now and then -inf */
    uint64_t r = rnd();
    if ((r & 0x3f) == 0) return -INFINITY;
    return (double)(-fabsf(rnd_float(-4, 17)));
}
static int test_cell_once(long trials) {
    long bad = 0, ties = 0;
    for (long i = 0; i < trials; i++) {
        uint64_t r = rnd();
        double lp_step = -fabs((double)rnd_float(-4, 2)) + (double)rnd_float(-45, -30);
        double lp_stay = -fabs((double)rnd_float(-5, 2)) + (double)rnd_float(-45, -30);
        double lp_skip = log(1e-10);
        double lpd = (double)(-fabsf(rnd_float(-1, 9)));
        double diag = neg_score(), up = neg_score(), left = neg_score();
        switch (r & 7) { /* force near-ties between the candidates: they decide `from` */
        case 0: case 1: { /* up such that u' is within a few float ulps of d' */
            double d = (diag + lp_step) + lpd;
            if (isfinite(d)) {
                float t = (float)(d - lp_stay - lpd);
                up = (double)f_from_bits(bits_from_f(t) + (uint32_t)((r >> 8) & 7) - 3u);
            }
            break; }
        case 2: case 3: { /* left such that l' is within a few ulps of max(d', u') */
            double d = (diag + lp_step) + lpd, u = (up + lp_stay) + lpd;
            double m = d > u ? d : u;
            if (isfinite(m)) {
                float t = (float)(m - lp_skip);
                left = (double)f_from_bits(bits_from_f(t) + (uint32_t)((r >> 8) & 7) - 3u);
            }
            break; }
        case 4: if ((r >> 8) & 1) diag = -INFINITY; if ((r >> 9) & 1) up = -INFINITY; if ((r >> 10) & 1) left = -INFINITY; break;
        default: break;
        }
        double s0, s1;
        int f0, f1;
        cell_ref(lpd, up, left, diag, lp_step, lp_stay, lp_skip, &s0, &f0);
        cell_once(lpd, up, left, diag, lp_step, lp_stay, lp_skip, &s1, &f1);
        {
            volatile float sd = (float)((diag + lp_step) + lpd), su = (float)((up + lp_stay) + lpd), sl = (float)(left + lp_skip);
            if (sd == su || su == sl || sd == sl) ties++;
        }
        if (bits_from_d(s0) != bits_from_d(s1) || f0 != f1) {
            if (bad < 10) fprintf(stderr, "cell mismatch: diag=%a up=%a left=%a lp=%a: ref (%a,%d) once (%a,%d)\n", diag, up, left, lpd, s0, f0, s1, f1);
            bad++;
        }
    }
    printf("cell (one rounding): %ld trials, %ld with tied rounded candidates, %ld mismatches\n", trials, ties, bad);
    return bad != 0;
}

int main(int argc, char** argv) {
    long m = argc > 1 ? atol(argv[1]) : 100;
    int rc = 0;
    rc |= test_emission(m * 1000000L);
    rc |= test_rounding(m * 1000000L);
    rc |= test_cell_once(m * 1000000L);
    return rc;
}
