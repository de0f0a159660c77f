// ecmc_math.cuh -- device arithmetic of the ECMC hot path (sm_100a, fp64).
//
// Everything here is written for the GPU: closed forms with the minimum number of fp64 divisions / square
// roots / cube roots, branch-light so that the lanes of a warp (= candidates of one chain) stay converged.
// It is NOT a transcription of the reference's Python: the algebra is regrouped (powers by multiplication,
// 1/6-th power by rcbrt, sphere points by their exact energy), results agree with the reference to ~1e-15
// relative (tests/test_gpu_potentials.py, tolerance 1e-12 as BASELINE.json states).
// Reference formulae are cited per function (paths relative to the reference checkout).
#pragma once

#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

#include "../../include/ecmc.h"
#include "ecmc_log_table.cuh"

#define ECMC_HD __host__ __device__ __forceinline__
#define ECMC_D __device__ __forceinline__

namespace ecmc {

// ---------------------------------------------------------------------------------------------------------
// Random stream: Philox4x32-10, key = (stream, seed), counter = (event_lo, event_hi, slot, block).
// Doubles are assembled like CPython's random(): (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53.
// ---------------------------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t w[4];
};

// One round costs two 32 x 32 -> 64 bit multiplications (IMAD.WIDE.U32: both halves of a product from one
// instruction), two three-input XORs and the two key increments.
ECMC_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; round++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        c1 = (uint32_t)p1;
        c2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 out;
    out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
    return out;
}

ECMC_HD double words_to_double(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    // the two integers written into the mantissa of 2^52 (exact) instead of two integer-to-double conversions:
    // (2^52 + B) 2^-53 - 1/2 = B 2^-53 and ((2^52 + A) - 2^52) 2^-27 + B 2^-53, every step exact
    const double low = fma(__hiloint2double(0x43300000, (int)(b >> 6)), 1.0 / 9007199254740992.0, -0.5);
    const double high = __hiloint2double(0x43300000, (int)(a >> 5)) - 4503599627370496.0;
    return fma(high, 1.0 / 134217728.0, low);
#else
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
#endif
}

struct StreamKey {
    uint32_t seed, stream;
    uint64_t event;
};

ECMC_HD Philox4 stream_block(const StreamKey &k, uint32_t slot, uint32_t block) {
    return philox4x32_10((uint32_t)k.event, (uint32_t)(k.event >> 32), slot, block, k.stream, k.seed);
}
// double `index` of a slot (two per block)
ECMC_HD double stream_double(const StreamKey &k, uint32_t slot, uint32_t index) {
    const Philox4 b = stream_block(k, slot, index >> 1);
    return (index & 1) ? words_to_double(b.w[2], b.w[3]) : words_to_double(b.w[0], b.w[1]);
}
ECMC_HD uint32_t stream_word(const StreamKey &k, uint32_t slot, uint32_t index) {
    const Philox4 b = stream_block(k, slot, index >> 2);
    return b.w[index & 3];
}
// CPython's Random._randbelow_with_getrandbits(n) with getrandbits(k) = next word >> (32 - k)
ECMC_HD uint32_t stream_randbelow(const StreamKey &k, uint32_t slot, uint32_t n) {
    int bits = 0;
    while ((n >> bits) != 0) bits++;
    uint32_t index = 0;
    for (;;) {
        const Philox4 b = stream_block(k, slot, index >> 2);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t r = b.w[j] >> (32 - bits);
            if (r < n) return r;
        }
        index += 4;
    }
}
// the same, continuing in the word stream of the slot at `index` (successive randint calls of one handler)
ECMC_HD uint32_t stream_randbelow_from(const StreamKey &k, uint32_t slot, uint32_t n, uint32_t &index) {
    int bits = 0;
    while ((n >> bits) != 0) bits++;
    for (;;) {
        const uint32_t r = stream_word(k, slot, index++) >> (32 - bits);
        if (r < n) return r;
    }
}
// random.expovariate(beta) = -log(1 - u) / beta
ECMC_D double expovariate(double u, double beta) { return -log(1.0 - u) / beta; }

// Natural logarithm for the argument range of expovariate, x = 1 - u in [2^-53, 1]: normal, positive, finite, so the
// special-case handling of the library routine (zero, negative, denormal, infinity, NaN) is dropped.
// x = 2^k m with m in [sqrt(1/2), sqrt(2)); c = m rounded to 8 binary digits (one addition of 1.5 * 2^44 does it and
// leaves j = 256 c in the low mantissa bits); log m = log c + log1p(r), r = (m - c) / c with |r| < 2.8e-3, where
// 1 / c and log c come from a table of 183 nodes (one 16-byte load) and log1p is a polynomial of degree 6
// (truncation error r^6 / 7 < 1e-16 relative). m - c is exact, and for x near 1 the node is c = 1 with log c = 0, so
// the result keeps full relative accuracy where expovariate needs it (small u). No division, no branch: ~25
// instructions instead of ~75 for the classic (m - 1) / (m + 1) reduction.
__device__ const double2 kLogTable[ECMC_LOG_TABLE_ENTRIES] = {ECMC_LOG_TABLE_VALUES};

ECMC_D double log_unit_interval(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int carry = (hi + 0x95f64) & 0x100000;  // mantissa above sqrt(2): halve it, bump the exponent
    k += carry >> 20;
    const double m = __hiloint2double(hi | (carry ^ 0x3ff00000), lo);
    const double shifted = m + 26388279066624.0;  // 1.5 * 2^44: rounds m to a multiple of 2^-8
    const double c = shifted - 26388279066624.0;
    const double2 node = __ldg(&kLogTable[__double2loint(shifted) - ECMC_LOG_TABLE_FIRST]);
    const double r = (m - c) * node.x;
    double q = fma(r, -1.0 / 6.0, 0.2);
    q = fma(r, q, -0.25);
    q = fma(r, q, 1.0 / 3.0);
    q = fma(r, q, -0.5);
    const double p = fma(r * r, q, r);
    // k as a double without a conversion instruction: k + 2^52 + 2^31 in the mantissa of 2^52
    const double dk = __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)k)) - 4503601774854144.0;
    return fma(dk, 6.93147180369123816490e-01, node.y + fma(dk, 1.90821492927058770002e-10, p));
}

// ---------------------------------------------------------------------------------------------------------
// Time as (quotient, remainder), jellyfysh/base/time.py. Time displacements here are >= 0.
// ---------------------------------------------------------------------------------------------------------
struct Time {
    double q, r;
};
ECMC_HD Time time_inf() {
    Time t;
    t.q = INFINITY; t.r = INFINITY;
    return t;
}
// Time.__add__ (time.py:115-133): divmod(r + dt, 1.0); for r + dt >= 0 this is floor / exact remainder.
ECMC_HD Time time_add(Time t, double dt) {
    Time out;
    if (isinf(dt) || isnan(dt)) {
        out.q = dt; out.r = dt;
        return out;
    }
    const double x = t.r + dt;
    const double fl = floor(x);
    out.q = t.q + fl;
    out.r = x - fl;
    return out;
}
// Time.__sub__ (time.py:135-149), same association
ECMC_HD double time_sub(Time a, Time b) { return a.q - b.q + a.r - b.r; }
// Time.__lt__ (time.py:167-182) == heap.c:176-178
ECMC_HD bool time_lt(Time a, Time b) { return a.q < b.q || (a.q == b.q && a.r < b.r); }

// ---------------------------------------------------------------------------------------------------------
// Periodic boundaries of the hypercubic box, jellyfysh/setting/hypercubic_setting.py:117,172.
// Python's float % has the sign of the divisor.
// ---------------------------------------------------------------------------------------------------------
// the general case of Python's float %, out of line: the hot paths only ever see arguments in [-L, 2L), and an
// inlined fmod is a hundred instructions at every call site
__host__ __device__ __noinline__ inline double py_mod_general(double x, double L) {
    double m = fmod(x, L);
    if (m != 0.0) {
        if (m < 0.0) m += L;
    } else {
        m = 0.0;
    }
    return m;
}
ECMC_HD double py_mod(double x, double L) {
    // x % L of Python for L > 0. In [-L, 2L) -- every case on the hot path -- fmod reduces to at most one exact
    // subtraction / one rounded addition, identical to CPython's float_rem; fmod itself is a long loop on the GPU.
    if (x >= 0.0) {
        if (x < L) return x;
        if (x < 2.0 * L) return x - L;
    } else if (x >= -L) {
        const double m = x + L;      // fmod(x, L) = x, then += L
        return m;
    }
    return py_mod_general(x, L);
}
ECMC_HD double correct_separation_entry(double s, double L, double half) { return py_mod(s + half, L) - half; }
// the same for two positions inside the box, |s| < L: s + L/2 lies in (-L/2, 3L/2) and the modulo is one select
ECMC_HD double correct_separation_in_box(double s, double L, double half) {
    const double x = s + half;
    const double shift = x < 0.0 ? L : (x >= L ? -L : 0.0);
    return (x + shift) - half;
}
ECMC_HD double correct_position_entry(double x, double L) { return py_mod(x, L); }

// ---------------------------------------------------------------------------------------------------------
// Potentials. Standard velocity: the active particle moves along +direction; `sd` is the separation
// component along it, `perp2` the squared norm of the other components (separation = target - active).
// All displacement functions return the DISPLACEMENT (length); callers divide by the speed.
// ---------------------------------------------------------------------------------------------------------
struct LennardJones {
    double k;        // prefactor
    double sigma;    // characteristic length
    double sigma2;   // sigma^2
    double r0sq;     // squared position of the minimum, 2^(1/3) sigma^2
    double four_over_k;
    double u_min;    // -k / 4
    double r_inflection_sq;  // squared distance of the largest attractive force, (26/7)^(1/3) sigma^2
    double force_max;        // that force, |U'| at the inflection point
};
ECMC_HD LennardJones make_lennard_jones(double prefactor, double characteristic_length) {
    LennardJones p;
    p.k = prefactor;
    p.sigma = characteristic_length;
    p.sigma2 = characteristic_length * characteristic_length;
    p.r0sq = 1.2599210498948731648 * p.sigma2;  // 2^(1/3)
    p.four_over_k = 4.0 / prefactor;
    p.u_min = -0.25 * prefactor;
    p.r_inflection_sq = pow(26.0 / 7.0, 1.0 / 3.0) * p.sigma2;
    p.force_max = prefactor / characteristic_length * (6.0 * pow(7.0 / 26.0, 7.0 / 6.0) - 12.0 * pow(7.0 / 26.0, 13.0 / 6.0));
    return p;
}
// U(r^2) = k [(s/r)^12 - (s/r)^6] (lennard_jones_potential.py:83-98) as k x3 (x3 - 1), x3 = (s^2/r^2)^3
ECMC_D double lj_energy(const LennardJones &p, double r2) {
    const double x = p.sigma2 / r2;
    const double x3 = x * x * x;
    return p.k * x3 * (x3 - 1.0);
}
// squared radius on the branch outside (inner = false) / inside (inner = true) of the minimum with energy u:
// (s/r)^6 = (1 -+ sqrt(1 + 4u/k)) / 2  (lennard_jones_potential.py:100-135), r^2 = s^2 / cbrt((s/r)^6)
ECMC_D double lj_radius_sq(const LennardJones &p, double u, bool inner) {
    const double root = sqrt(fmax(fma(u, p.four_over_k, 1.0), 0.0));
    const double x3 = 0.5 * (inner ? 1.0 + root : 1.0 - root);
    return p.sigma2 * rcbrt(x3);
}
// derivative of the pair energy with respect to the active particle's coordinate along the direction of motion
// (lennard_jones_potential.py:63-81): 12 sd k s^12 / r^14 - 6 sd k s^6 / r^8 = sd k x3 (12 x3 - 6) / r^2
ECMC_D double lj_derivative(const LennardJones &p, double sd, double perp2) {
    const double r2 = fma(sd, sd, perp2);
    const double inv = 1.0 / r2;
    const double x = p.sigma2 * inv;
    const double x3 = x * x * x;
    return sd * p.k * x3 * (12.0 * x3 - 6.0) * inv;
}
// Event-rate inversion of MexicanHatPotential.standard_velocity_displacement (potential/abstracts.py:336-530)
// for the Lennard-Jones energy. The four reference branches (outside/inside the minimum sphere x in
// front/behind) collapse to at most two uphill stretches along the straight line of the active particle:
//   (1) inside the sphere while approaching (sd from min(sd, h) down to 0), if the line hits the sphere;
//   (2) outside the sphere while receding (sd from -h, or from min(sd, 0) if the sphere is missed, to -inf).
// h = sqrt(r0^2 - perp2) is the half chord. Energies at sphere points are the exact minimum -k/4.
// Written without divergent branches -- the lanes of a warp hold different targets, so every branch any lane takes
// costs the whole warp: both energies come from ONE division, the case analysis is a chain of selects, and every
// lane runs the same sqrt -> rcbrt -> sqrt sequence once.
ECMC_D double lj_displacement(const LennardJones &p, double sd, double perp2, double du) {
    const double r2 = fma(sd, sd, perp2);
    const double pc = fmax(perp2, 1.0e-150);     // head-on approach: the barrier is infinite
    const double inv = 1.0 / (r2 * pc);
    const double xr = p.sigma2 * (inv * pc), xp = p.sigma2 * (inv * r2);
    const double xr3 = xr * xr * xr, xp3 = xp * xp * xp;
    const double e_r = p.k * xr3 * (xr3 - 1.0);  // U at the current separation
    const double e_p = p.k * xp3 * (xp3 - 1.0);  // U at the closest approach of the straight line
    const bool approaching = sd > 0.0;
    const bool hits = perp2 < p.r0sq;            // the line of motion crosses the minimum sphere
    const bool inside = r2 < p.r0sq;
    const bool climbs = approaching && hits;     // stretch (1) exists
    const double u1 = inside ? e_r : p.u_min;    // energy where stretch (1) starts
    const double barrier = e_p - u1;
    const bool inner = climbs && du < barrier;   // the event happens on stretch (1)
    // energy at which the event happens: on stretch (1), or on stretch (2) started at the sphere / the closest
    // approach / the current point
    const double u_start = climbs ? p.u_min : (approaching ? e_p : (inside ? p.u_min : e_r));
    const double u_event = inner ? u1 + du : u_start + (climbs ? du - barrier : du);
    const double root = sqrt(fmax(fma(u_event, p.four_over_k, 1.0), 0.0));
    const double x3 = 0.5 * (inner ? 1.0 + root : 1.0 - root);
    const double rn2 = p.sigma2 * rcbrt(x3);
    const double s = sqrt(rn2 - perp2);
    if (!inner && u_event >= 0.0) return INFINITY;  // escapes the attractive tail
    return inner ? sd - s : sd + s;
}

// Cheap exclusion test for lj_displacement: a candidate that does not climb the repulsive core (stretch (1)) starts
// stretch (2) at an energy u_start < 0 and escapes to infinity if its potential change exceeds -u_start. With a lower
// bound of that potential change this needs one division and no logarithm / root. The margin keeps the test
// strictly on the safe side of rounding: anything close to the threshold goes through the full computation.
ECMC_D bool lj_certainly_dead(const LennardJones &p, double sd, double perp2, double du_lower) {
    const bool approaching = sd > 0.0;
    if (approaching && perp2 < p.r0sq) return false;  // hits the minimum sphere: climbs the core
    const double r2 = approaching ? perp2 : fma(sd, sd, perp2);  // where stretch (2) starts
    if (r2 < p.r0sq) return du_lower > -p.u_min * (1.0 + 1.0e-9);
    const double x = p.sigma2 / r2;
    const double x3 = x * x * x;
    const double u_start = p.k * x3 * (x3 - 1.0);
    return du_lower > -u_start * (1.0 + 1.0e-9);
}

// Can the Lennard-Jones candidate fire within the next `d` of displacement? `du_lower` <= the potential change it will
// draw. False only if it certainly cannot: the stretch [0, d] holds neither the closest approach nor a crossing of the
// minimum sphere, so the energy is monotonic along it and rises by U(end) - U(now), which one division gives; anything
// else, and anything within the rounding margin, is left to the full computation.
ECMC_D bool lj_may_fire_within(const LennardJones &p, double sd, double perp2, double d, double du_lower) {
    if (!(d < INFINITY)) return true;
    if (sd > 0.0 && !(d < sd)) return true;          // the closest approach lies inside the stretch
    const double e = sd - d;
    const double r2_now = fma(sd, sd, perp2), r2_end = fma(e, e, perp2);
    if ((r2_now < p.r0sq) != (r2_end < p.r0sq)) return true;  // crosses the minimum sphere
    const double inv = 1.0 / (r2_now * r2_end);
    const double x_now = p.sigma2 * (inv * r2_end), x_end = p.sigma2 * (inv * r2_now);
    const double x3_now = x_now * x_now * x_now, x3_end = x_end * x_end * x_end;
    const double rise = p.k * (x3_end * (x3_end - 1.0) - x3_now * (x3_now - 1.0));
    return !(rise + 1.0e-12 < du_lower * (1.0 - 1.0e-9));
}

// ---- hard sphere / hard dipole, general velocity (hard_sphere_potential.py:65-99, hard_dipole_potential.py:75-114)
// Grazing collisions make the square-root term cancel to ~0, where one ulp of its inputs decides between a hit
// and a miss. These few operations therefore follow the reference operation by operation: products and sums
// rounded separately (no fma contraction) and Python's builtin sum() -- Neumaier-compensated since CPython 3.12 --
// for the squared norms and the dot product (base/vectors.py:60, :122).
ECMC_D double pysum3(double a, double b, double c) {
    double f = a, comp = 0.0;
    double t = __dadd_rn(f, b);
    comp = __dadd_rn(comp, fabs(f) >= fabs(b) ? __dadd_rn(__dsub_rn(f, t), b) : __dadd_rn(__dsub_rn(b, t), f));
    f = t;
    t = __dadd_rn(f, c);
    comp = __dadd_rn(comp, fabs(f) >= fabs(c) ? __dadd_rn(__dsub_rn(f, t), c) : __dadd_rn(__dsub_rn(c, t), f));
    f = t;
    if (comp != 0.0 && isfinite(comp)) f = __dadd_rn(f, comp);
    return f;
}
ECMC_D double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return pysum3(__dmul_rn(ax, bx), __dmul_rn(ay, by), __dmul_rn(az, bz));
}
// returns a TIME. vv = |v|^2, vs = v.s, ss = |s|^2
ECMC_D double hard_sphere_time(double radius, double vv, double vs, double ss) {
    const double d2 = __dmul_rn(__dmul_rn(4.0, radius), radius);
    const double term = __dsub_rn(__dmul_rn(vs, vs), __dmul_rn(vv, __dsub_rn(ss, d2)));
    return (term >= 0.0 && vs >= 0.0) ? __dsub_rn(vs, sqrt(term)) / vv : INFINITY;
}
ECMC_D double hard_dipole_time(double min_sep, double max_sep, double vv, double vs, double ss) {
    if (vs >= 0.0) {
        const double term = __dsub_rn(__dmul_rn(vs, vs), __dmul_rn(vv, __dsub_rn(ss, __dmul_rn(min_sep, min_sep))));
        if (term >= 0.0) return __dsub_rn(vs, sqrt(term)) / vv;
    }
    const double term = __dsub_rn(__dmul_rn(vs, vs), __dmul_rn(vv, __dsub_rn(ss, __dmul_rn(max_sep, max_sep))));
    return __dadd_rn(vs, sqrt(term)) / vv;
}

// ---- inverse power: U = c k / r^p (inverse_power_potential.py) -------------------------------------------
// Small integer powers (every shipped configuration) avoid the generic pow: r^p by multiplication (times one
// square root for odd p) and the inverse x^(2/p) by sqrt / cbrt. That is both much cheaper on the fp64 pipe and,
// for p <= 2, identical to the reference's libm results, which the "barely can / cannot escape" known-answer
// tests of the reference need: there U + dU cancels to ~100 ulp and every rounding of U shows in the result.
// For the same reason r2 and perp2 are formed by the caller like the reference forms them (ip_squares).
struct InversePower {
    double power, k;
    int int_power;  // the power if it is an integer in [1, 32], else 0
    int pad;
};
ECMC_HD InversePower make_inverse_power(double power, double prefactor) {
    InversePower p;
    p.power = power;
    p.k = prefactor;
    p.int_power = (power >= 1.0 && power <= 32.0 && power == (double)(int)power) ? (int)power : 0;
    p.pad = 0;
    return p;
}
// r2^(power / 2)
ECMC_D double ip_pow_half(const InversePower &p, double r2) {
    if (p.int_power == 0) return pow(r2, 0.5 * p.power);
    double result = (p.int_power & 1) ? sqrt(r2) : 1.0;
    double base = r2;
    for (int e = p.int_power >> 1; e > 0; e >>= 1) {
        if (e & 1) result *= base;
        base *= base;
    }
    return result;
}
// x^(2 / power)
ECMC_D double ip_root(const InversePower &p, double x) {
    switch (p.int_power) {
    case 1: return x * x;
    case 2: return x;
    case 3: { const double c = cbrt(x); return c * c; }
    case 4: return sqrt(x);
    case 6: return cbrt(x);
    case 8: return sqrt(sqrt(x));
    case 12: return cbrt(sqrt(x));
    default: return pow(x, 2.0 / p.power);
    }
}
// norm_sq (base/vectors.py:60, Python's compensated sum) and the squared norm of the components other than the
// direction of motion (base/vectors.py:181, a plain two-term sum)
ECMC_D void ip_squares(int dir, double sx, double sy, double sz, double &r2, double &perp2) {
    r2 = dot3(sx, sy, sz, sx, sy, sz);
    const double xx = __dmul_rn(sx, sx), yy = __dmul_rn(sy, sy), zz = __dmul_rn(sz, sz);
    perp2 = dir == 0 ? __dadd_rn(yy, zz) : (dir == 1 ? __dadd_rn(xx, zz) : __dadd_rn(xx, yy));
}
ECMC_D double ip_energy(const InversePower &p, double kc, double r2) { return kc / ip_pow_half(p, r2); }
// derivative, inverse_power_potential.py:71-94: power * sd / r^(power + 2) * k * c1 * c2
ECMC_D double ip_derivative(const InversePower &p, double sd, double r2, double c1, double c2) {
    return p.power * sd / (ip_pow_half(p, r2) * r2) * p.k * c1 * c2;
}
// displacement, inverse_power_potential.py:96-179
ECMC_D double ip_displacement(const InversePower &p, double sd, double perp2, double r2, double c1, double c2, double du) {
    const double cc = c1 * c2;
    const double kc = cc * p.k;
    if (p.k * cc > 0.0) {
        if (sd <= 0.0) return INFINITY;
        const double u_max = ip_energy(p, kc, perp2);
        const double u_now = ip_energy(p, kc, r2);
        if (du < __dsub_rn(u_max, u_now)) {
            const double rn2 = ip_root(p, kc / __dadd_rn(u_now, du));
            return __dsub_rn(sd, sqrt(__dsub_rn(rn2, perp2)));
        }
        return INFINITY;
    }
    double done = 0.0;
    double start2 = r2;
    if (sd > 0.0) {
        done = sd;
        sd = 0.0;
        start2 = perp2;
    }
    const double u_now = ip_energy(p, kc, start2);
    const double u_end = __dadd_rn(u_now, du);
    if (u_end >= 0.0) return INFINITY;
    const double rn2 = ip_root(p, kc / u_end);
    return __dadd_rn(done, __dadd_rn(sd, sqrt(__dsub_rn(rn2, perp2))));
}

// ---- displaced even power: U = k (r - r0)^p (displaced_even_power_potential.py) ----------------------------
struct DisplacedEvenPower {
    double k, r0, power;
};
// x^power and x^(1/power); SQUARE: the harmonic bond (power = 2, every shipped configuration), known at compile time,
// needs no generic pow (a few hundred instructions per call site)
template <bool SQUARE>
ECMC_D double dep_pow(const DisplacedEvenPower &p, double x) { return (SQUARE || p.power == 2.0) ? x * x : pow(x, p.power); }
template <>
ECMC_D double dep_pow<true>(const DisplacedEvenPower &, double x) { return x * x; }
template <bool SQUARE>
ECMC_D double dep_root(const DisplacedEvenPower &p, double x) { return p.power == 2.0 ? sqrt(x) : pow(x, 1.0 / p.power); }
template <>
ECMC_D double dep_root<true>(const DisplacedEvenPower &, double x) { return sqrt(x); }
template <bool SQUARE = false>
ECMC_D double dep_energy(const DisplacedEvenPower &p, double r2) { return p.k * dep_pow<SQUARE>(p, sqrt(r2) - p.r0); }
ECMC_D double dep_derivative(const DisplacedEvenPower &p, double sd, double perp2) {
    const double r = sqrt(fma(sd, sd, perp2));
    const double d = r - p.r0;
    return -p.power * p.k * (p.power == 2.0 ? d : pow(d, p.power - 1.0)) * sd / r;
}
// same two-stretch structure as lj_displacement; the outside branch never escapes (U -> +inf)
template <bool SQUARE = false>
ECMC_D double dep_displacement(const DisplacedEvenPower &p, double sd, double perp2, double du) {
    const double r0sq = p.r0 * p.r0;
    const double r2 = fma(sd, sd, perp2);
    const bool hits = perp2 < r0sq;
    const bool inside = r2 < r0sq;
    double u_start;
    if (sd > 0.0 && hits) {
        const double u1 = inside ? dep_energy<SQUARE>(p, r2) : 0.0;
        const double u_max = dep_energy<SQUARE>(p, perp2);
        const double barrier = u_max - u1;
        if (du < barrier) {
            const double rn = p.r0 - dep_root<SQUARE>(p, (u1 + du) / p.k);
            return sd - sqrt(rn * rn - perp2);
        }
        du -= barrier;
        u_start = 0.0;
    } else if (sd > 0.0) {
        u_start = dep_energy<SQUARE>(p, perp2);
    } else {
        u_start = inside ? 0.0 : dep_energy<SQUARE>(p, r2);
    }
    const double rn = p.r0 + dep_root<SQUARE>(p, (u_start + du) / p.k);
    return sd + sqrt(rn * rn - perp2);
}

// ---- inverse-power Coulomb bound (inverse_power_coulomb_bounding_potential.c:53-139) ----------------------
// sx is the component along the direction of motion, p2 = sy^2 + sz^2, kc = prefactor c1 c2.
ECMC_D double ipcb_derivative(double kc, double sx, double p2) {
    const double r2 = fma(sx, sx, p2);
    const double inv_r = rsqrt(r2);
    return kc * sx * (inv_r * inv_r * inv_r);
}
ECMC_D double ipcb_displacement(double kc, double sx, double p2, double du, double L) {
    const double half = 0.5 * L;
    const double u_now = kc * rsqrt(fma(sx, sx, p2));
    const double u_zero = kc * rsqrt(p2);
    const double u_half = kc * rsqrt(fma(half, half, p2));
    const double per_length = fabs(u_zero - u_half);
    const double laps = floor(du / per_length);
    double disp = laps * L;
    du = fma(-laps, per_length, du);               // fmod(du, per_length) for du >= 0
    if (du < 0.0) du = 0.0;
    double start = u_now;
    if (kc > 0.0) {
        if (sx <= 0.0) {
            disp += half + sx;
            sx = half;
            start = u_half;
        } else if (du >= u_zero - u_now) {
            du -= (u_zero - u_now);
            disp += sx + half;
            sx = half;
            start = u_half;
        }
        const double rn = kc / (start + du);
        return disp + (sx - sqrt(fma(rn, rn, -p2)));
    }
    if (sx > 0.0) {
        disp += sx;
        sx = 0.0;
        start = u_zero;
    } else if (du >= u_half - u_now) {
        du -= (u_half - u_now);
        disp += sx + L;
        sx = 0.0;
        start = u_zero;
    }
    const double rn = kc / (start + du);
    return disp + (sx + sqrt(fma(rn, rn, -p2)));
}

// ---- merged-image Coulomb (merged_image_coulomb_potential.c:77-274) -----------------------------------------
// The Ewald sum is spread over the lanes of a warp: every lane takes a share of the real-space images and of the
// Fourier modes (flat term tables built on the host), then the partial sums are added by shuffles.
// Must be called by all 32 lanes of a converged warp with identical arguments.
struct MergedImageCoulomb {
    double prefactor;
    double alpha_over_length;      // alpha / L
    double alpha_over_length_sq;
    double two_alpha_root_pi;      // 2 alpha / (L sqrt(pi))
    double length;
    double two_pi_over_length;
    int n_images;                  // real-space images (i, j, k), packed as int8 x 3 in an int
    int n_modes;                   // Fourier modes (i >= 1, j >= 0, k >= 0)
    int fourier_cutoff;
    const int *images;             // [n_images]  i | j << 8 | k << 16 (signed bytes)
    const int *modes;              // [n_modes]   i | j << 8 | k << 16
    const double *coefficients;    // [n_modes]
};

ECMC_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// derivative along x of sum_n 1/|s + nL| for s = (sx, sy, sz) (already permuted so that x is the direction of
// motion); `trig` is per-warp shared scratch of 3 * 2 * (fourier_cutoff + 1) doubles.
ECMC_D double mic_derivative_warp(const MergedImageCoulomb &p, double sx, double sy, double sz, double *trig,
                                  int lane) {
    double acc = 0.0;
    for (int t = lane; t < p.n_images; t += 32) {
        const int code = __ldg(p.images + t);
        const double vx = fma((double)(signed char)(code & 0xff), p.length, sx);
        const double vy = fma((double)(signed char)((code >> 8) & 0xff), p.length, sy);
        const double vz = fma((double)(signed char)((code >> 16) & 0xff), p.length, sz);
        const double r2 = fma(vx, vx, fma(vy, vy, vz * vz));
        const double inv_r = rsqrt(r2);
        const double r = r2 * inv_r;
        acc += vx * fma(p.two_alpha_root_pi, exp(-p.alpha_over_length_sq * r2), erfc(p.alpha_over_length * r) * inv_r)
               * (inv_r * inv_r);
    }
    // sin / cos of m * 2 pi s / L for m = 0..fc on the three axes: one sincos per lane
    const int per_axis = p.fourier_cutoff + 1;
    __syncwarp();
    for (int t = lane; t < 3 * per_axis; t += 32) {
        const int axis = t / per_axis, m = t - axis * per_axis;
        const double s = axis == 0 ? sx : (axis == 1 ? sy : sz);
        double sn, cs;
        sincos(p.two_pi_over_length * s * (double)m, &sn, &cs);
        trig[2 * t] = cs;
        trig[2 * t + 1] = sn;
    }
    __syncwarp();
    for (int t = lane; t < p.n_modes; t += 32) {
        const int code = __ldg(p.modes + t);
        const int i = code & 0xff, j = (code >> 8) & 0xff, k = (code >> 16) & 0xff;
        acc = fma(__ldg(p.coefficients + t) * trig[2 * i + 1], trig[2 * (per_axis + j)] * trig[2 * (2 * per_axis + k)],
                  acc);
    }
    __syncwarp();
    return warp_sum(acc);
}

// Three of these sums side by side (the active leaf of a molecule against the three leaves of a target molecule): the 3 x 33
// image terms, the 3 x 21 sine / cosine pairs and the modes are spread over the lanes TOGETHER -- four passes of exp / erfc
// instead of six (the 33rd image of a single sum costs a pass of its own), two of sincos instead of three, and one load of
// every mode for three products. Same terms as mic_derivative_warp; only the order of the additions differs (1e-16).
// `trig`: per-warp scratch of 3 * 3 * 2 * (fourier_cutoff + 1) doubles.
ECMC_D void mic_derivative_warp3(const MergedImageCoulomb &p, double x0, double y0, double z0, double x1, double y1,
                                 double z1, double x2, double y2, double z2, double *trig, int lane, double &out0,
                                 double &out1, double &out2) {
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    const int n = p.n_images;
    for (int f = lane; f < 3 * n; f += 32) {
        const int q = f >= 2 * n ? 2 : (f >= n ? 1 : 0);
        const int code = __ldg(p.images + (f - q * n));
        const double vx = fma((double)(signed char)(code & 0xff), p.length, q == 0 ? x0 : (q == 1 ? x1 : x2));
        const double vy = fma((double)(signed char)((code >> 8) & 0xff), p.length, q == 0 ? y0 : (q == 1 ? y1 : y2));
        const double vz = fma((double)(signed char)((code >> 16) & 0xff), p.length, q == 0 ? z0 : (q == 1 ? z1 : z2));
        const double r2 = fma(vx, vx, fma(vy, vy, vz * vz));
        const double inv_r = rsqrt(r2);
        const double r = r2 * inv_r;
        const double term = vx * fma(p.two_alpha_root_pi, exp(-p.alpha_over_length_sq * r2), erfc(p.alpha_over_length * r) * inv_r)
                            * (inv_r * inv_r);
        if (q == 0) acc0 += term; else if (q == 1) acc1 += term; else acc2 += term;
    }
    const int per_axis = p.fourier_cutoff + 1, per_sum = 3 * per_axis;
    __syncwarp();
    for (int f = lane; f < 3 * per_sum; f += 32) {
        const int q = f / per_sum, t = f - q * per_sum;
        const int axis = t / per_axis, m = t - axis * per_axis;
        const double s = q == 0 ? (axis == 0 ? x0 : (axis == 1 ? y0 : z0))
                                : (q == 1 ? (axis == 0 ? x1 : (axis == 1 ? y1 : z1)) : (axis == 0 ? x2 : (axis == 1 ? y2 : z2)));
        double sn, cs;
        sincos(p.two_pi_over_length * s * (double)m, &sn, &cs);
        trig[2 * f] = cs;
        trig[2 * f + 1] = sn;
    }
    __syncwarp();
    const double *trig1 = trig + 2 * per_sum, *trig2 = trig + 4 * per_sum;
    for (int t = lane; t < p.n_modes; t += 32) {
        const int code = __ldg(p.modes + t);
        const int i = 2 * (code & 0xff) + 1, j = 2 * (per_axis + ((code >> 8) & 0xff)), k = 2 * (2 * per_axis + ((code >> 16) & 0xff));
        const double coefficient = __ldg(p.coefficients + t);
        acc0 = fma(coefficient * trig[i], trig[j] * trig[k], acc0);
        acc1 = fma(coefficient * trig1[i], trig1[j] * trig1[k], acc1);
        acc2 = fma(coefficient * trig2[i], trig2[j] * trig2[k], acc2);
    }
    __syncwarp();
    out0 = warp_sum(acc0);
    out1 = warp_sum(acc1);
    out2 = warp_sum(acc2);
}

}  // namespace ecmc
