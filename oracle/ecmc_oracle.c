/* ecmc_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the JeLLyFysh algorithms on the hot path (next-event-time computation for the
 * active particle -> argmin -> lifting -> commit), written function by function after the reference's
 * Python / C sources, with the same operation order, Python float semantics (`%`, divmod, `**` = libm pow)
 * and libm calls. Paths cited below are relative to the reference checkout.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this file.
 * The product (jellyfysh_b200/, libecmc_b200.so) never links, imports or calls it.
 *
 * Parity pinning: tests/test_oracle_*.py check this file against (i) the known-answer constants of the
 * reference's own unit tests and (ii) golden vectors and whole-chain traces recorded from the running
 * reference (tests/golden/make_golden.py).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/ecmc.h"

#define ORC_API __attribute__((visibility("default")))
#define ORC_INF (1.0 / 0.0)

/* ================================================================================================== */
/* Python float semantics                                                                             */
/* ================================================================================================== */

/* CPython float_rem (Objects/floatobject.c): result has the sign of the divisor. Used by
 * jellyfysh/setting/hypercubic_setting.py:117,172. */
static double py_mod(double vx, double wx) {
    double mod = fmod(vx, wx);
    if (mod) {
        if ((wx < 0) != (mod < 0)) mod += wx;
    } else {
        mod = copysign(0.0, wx);
    }
    return mod;
}

/* CPython float_divmod. Used by jellyfysh/base/time.py:101,131. */
static void py_divmod(double vx, double wx, double *quotient, double *remainder) {
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    double floordiv;
    if (mod) {
        if ((wx < 0) != (mod < 0)) {
            mod += wx;
            div -= 1.0;
        }
    } else {
        mod = copysign(0.0, wx);
    }
    if (div) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        floordiv = copysign(0.0, vx / wx);
    }
    *quotient = floordiv;
    *remainder = mod;
}

/* ================================================================================================== */
/* base/time.py                                                                                       */
/* ================================================================================================== */
typedef struct { double q, r; } otime;

/* Time.__add__, jellyfysh/base/time.py:115-133 */
static otime time_add(otime t, double other) {
    otime out;
    if (!isinf(other)) {
        double add_q, new_r;
        py_divmod(t.r + other, 1.0, &add_q, &new_r);
        out.q = t.q + add_q;
        out.r = new_r;
    } else {
        out.q = other;
        out.r = other;
    }
    return out;
}
/* Time.__sub__, jellyfysh/base/time.py:135-149 */
static double time_sub(otime a, otime b) { return a.q - b.q + a.r - b.r; }
/* Time.__lt__, jellyfysh/base/time.py:167-182; same comparison as heap.c:176-178 */
static int time_lt(otime a, otime b) { return a.q < b.q || (a.q == b.q && a.r < b.r); }
/* Time.from_float, jellyfysh/base/time.py:87-101 */
static otime time_from_float(double t) {
    otime out;
    if (!isinf(t)) py_divmod(t, 1.0, &out.q, &out.r);
    else { out.q = t; out.r = t; }
    return out;
}

ORC_API void orc_time_add(double q, double r, double other, double *out_q, double *out_r) {
    otime t = {q, r};
    t = time_add(t, other);
    *out_q = t.q; *out_r = t.r;
}
ORC_API double orc_time_sub(double q1, double r1, double q2, double r2) {
    otime a = {q1, r1}, b = {q2, r2};
    return time_sub(a, b);
}
/* the order of the scheduler's heap, heap.c:176-178 / :217-233 (Time.__lt__, base/time.py:167-182) */
ORC_API int orc_time_lt(double q1, double r1, double q2, double r2) {
    otime a = {q1, r1}, b = {q2, r2};
    return time_lt(a, b);
}
ORC_API void orc_time_from_float(double t, double *out_q, double *out_r) {
    otime x = time_from_float(t);
    *out_q = x.q; *out_r = x.r;
}

/* ================================================================================================== */
/* setting/hypercubic_setting.py: periodic boundaries                                                 */
/* ================================================================================================== */
/* correct_position_entry, hypercubic_setting.py:117 */
static double correct_position_entry(double x, double L) { return py_mod(x, L); }
/* correct_separation_entry, hypercubic_setting.py:172 */
static double correct_separation_entry(double s, double L) {
    double half = L / 2.0;
    return py_mod(s + half, L) - half;
}
/* separation_vector, hypercubic_setting.py:138-140 */
static void separation_vector(const double *ref, const double *target, int D, double L, double *sep) {
    for (int d = 0; d < D; d++) sep[d] = correct_separation_entry(target[d] - ref[d], L);
}
ORC_API double orc_correct_position_entry(double x, double L) { return correct_position_entry(x, L); }
ORC_API double orc_correct_separation_entry(double s, double L) { return correct_separation_entry(s, L); }

/* ================================================================================================== */
/* base/vectors.py                                                                                    */
/* ================================================================================================== */
/* Python's builtin sum() over floats. CPython >= 3.12 (the interpreter the reference runs under in this
 * container and on the GPU box) uses Neumaier compensated summation (Python/bltinmodule.c, builtin_sum_impl);
 * PyPy and CPython < 3.12 add naively. The start value is int 0, so the first float is taken as is.
 * Mode 1 (default) = CPython >= 3.12, mode 0 = naive. */
static int g_sum_mode = 1;
ORC_API void orc_set_sum_mode(int compensated) { g_sum_mode = compensated; }
typedef struct { double f, c; int n; } pysum;
static void pysum_init(pysum *s) { s->f = 0.0; s->c = 0.0; s->n = 0; }
static void pysum_add(pysum *s, double x) {
    if (s->n++ == 0) { s->f = 0 + x; return; }
    if (!g_sum_mode) { s->f = s->f + x; return; }
    double t = s->f + x;
    if (fabs(s->f) >= fabs(x)) s->c += (s->f - t) + x;
    else s->c += (x - t) + s->f;
    s->f = t;
}
static double pysum_result(const pysum *s) {
    if (s->n == 0) return 0.0;
    double f = s->f;
    if (g_sum_mode && s->c && isfinite(s->c)) f += s->c;
    return f;
}
/* norm_sq, base/vectors.py:60 */
static double v_norm_sq(const double *v, int D) {
    pysum s;
    pysum_init(&s);
    for (int d = 0; d < D; d++) pysum_add(&s, v[d] * v[d]);
    return pysum_result(&s);
}
/* norm: `** 0.5` is libm pow, base/vectors.py:43 */
static double v_norm(const double *v, int D) { return pow(v_norm_sq(v, D), 0.5); }

/* math.sqrt raises ValueError on a negative argument; the reference uses this as control flow
 * (jellyfysh/potential/abstracts.py:440-454). *err is set instead. */
static double checked_sqrt(double x, int *err) {
    if (x < 0.0) { *err = 1; return NAN; }
    return sqrt(x);
}
static double other_components_sq(const double *v, int D, int dir) {
    pysum s;
    pysum_init(&s);
    for (int d = 0; d < D; d++)
        if (d != dir) pysum_add(&s, pow(v[d], 2.0));
    return pysum_result(&s);
}
/* displacement_until_new_norm_sq_component_positive, base/vectors.py:153-182 */
static double disp_new_norm_sq_positive(const double *v, int D, double new_norm_sq, int dir, int *err) {
    return v[dir] - checked_sqrt(new_norm_sq - other_components_sq(v, D, dir), err);
}
/* displacement_until_new_norm_sq_component_negative, base/vectors.py:185-214 */
static double disp_new_norm_sq_negative(const double *v, int D, double new_norm_sq, int dir, int *err) {
    return v[dir] + checked_sqrt(new_norm_sq - other_components_sq(v, D, dir), err);
}

/* ================================================================================================== */
/* potential/inverse_power_potential.py                                                               */
/* ================================================================================================== */
typedef struct {
    double power, prefactor;
    double two_over_power, power_over_two, power_plus_two;
} inverse_power;

static inverse_power ip_make(double power, double prefactor) {
    inverse_power p;
    p.power = power; p.prefactor = prefactor;
    p.two_over_power = 2.0 / power;       /* inverse_power_potential.py:66 */
    p.power_over_two = power / 2.0;       /* :67 */
    p.power_plus_two = power + 2;         /* :68 */
    return p;
}
/* standard_velocity_derivative, inverse_power_potential.py:71-94 */
static double ip_derivative(const inverse_power *p, int dir, const double *sep, int D, double c1, double c2) {
    return p->power * sep[dir] / pow(v_norm(sep, D), p->power_plus_two) * p->prefactor * c1 * c2;
}
/* potential, inverse_power_potential.py:126-143 */
static double ip_potential(const inverse_power *p, double charge_product, const double *sep, int D) {
    return charge_product * p->prefactor / pow(v_norm_sq(sep, D), p->power_over_two);
}
/* _displacement_repulsive, inverse_power_potential.py:145-160 */
static double ip_displacement_repulsive(const inverse_power *p, int dir, double cp, double dU, double *sep, int D,
                                        int *err) {
    if (sep[dir] <= 0.0) return ORC_INF;
    double tmp[ECMC_MAX_DIM];
    memcpy(tmp, sep, sizeof(double) * D);
    tmp[dir] = 0.0;
    double maximum_potential = ip_potential(p, cp, tmp, D);
    double current_potential = ip_potential(p, cp, sep, D);
    if (dU < maximum_potential - current_potential) {
        double new_norm_sq = pow(cp * p->prefactor / (current_potential + dU), p->two_over_power);
        return disp_new_norm_sq_positive(sep, D, new_norm_sq, dir, err);
    }
    return ORC_INF;
}
/* _displacement_attractive, inverse_power_potential.py:162-179 */
static double ip_displacement_attractive(const inverse_power *p, int dir, double cp, double dU, double *sep, int D,
                                         int *err) {
    double current_displacement = 0.0;
    if (sep[dir] > 0.0) {
        current_displacement += sep[dir];
        sep[dir] = 0.0;
    }
    double current_potential = ip_potential(p, cp, sep, D);
    if (current_potential + dU >= 0.0) return ORC_INF;
    double new_norm_sq = pow(cp * p->prefactor / (current_potential + dU), p->two_over_power);
    current_displacement += disp_new_norm_sq_negative(sep, D, new_norm_sq, dir, err);
    return current_displacement;
}
/* standard_velocity_displacement, inverse_power_potential.py:96-124 */
static double ip_displacement(const inverse_power *p, int dir, double *sep, int D, double c1, double c2, double dU,
                              int *err) {
    double cp = c1 * c2;
    double prefactor_product = p->prefactor * cp;
    return prefactor_product > 0 ? ip_displacement_repulsive(p, dir, cp, dU, sep, D, err)
                                 : ip_displacement_attractive(p, dir, cp, dU, sep, D, err);
}

/* ================================================================================================== */
/* potential/abstracts.py: MexicanHatPotential; lennard_jones_potential.py; displaced_even_power...   */
/* ================================================================================================== */
typedef struct {
    int kind; /* ECMC_POT_LENNARD_JONES or ECMC_POT_DISPLACED_EVEN_POWER */
    double prefactor, equilibrium_separation, equilibrium_separation_squared;
    /* LJ */
    double characteristic_length;
    inverse_power six, twelve;
    /* displaced even power */
    double power, inverse_power_;
} mexican_hat;

/* LennardJonesPotential.__init__, lennard_jones_potential.py:42-61 */
static mexican_hat lj_make(double prefactor, double characteristic_length) {
    mexican_hat m;
    memset(&m, 0, sizeof(m));
    m.kind = ECMC_POT_LENNARD_JONES;
    m.prefactor = prefactor;
    m.equilibrium_separation = characteristic_length * pow(2.0, 1.0 / 6.0);
    m.equilibrium_separation_squared = pow(m.equilibrium_separation, 2.0); /* abstracts.py:323 */
    m.characteristic_length = characteristic_length;
    m.six = ip_make(6.0, -prefactor * pow(characteristic_length, 6.0));
    m.twelve = ip_make(12.0, prefactor * pow(characteristic_length, 12.0));
    return m;
}
/* DisplacedEvenPowerPotential.__init__, displaced_even_power_potential.py:45-71 */
static mexican_hat dep_make(double prefactor, double equilibrium_separation, double power) {
    mexican_hat m;
    memset(&m, 0, sizeof(m));
    m.kind = ECMC_POT_DISPLACED_EVEN_POWER;
    m.prefactor = prefactor;
    m.equilibrium_separation = equilibrium_separation;
    m.equilibrium_separation_squared = equilibrium_separation * equilibrium_separation;
    m.power = power;
    m.inverse_power_ = 1.0 / power;
    return m;
}
static double mh_potential(const mexican_hat *m, const double *sep, int D) {
    if (m->kind == ECMC_POT_LENNARD_JONES) {
        /* lennard_jones_potential.py:83-98 */
        return ip_potential(&m->six, 1.0, sep, D) + ip_potential(&m->twelve, 1.0, sep, D);
    }
    /* displaced_even_power_potential.py:96-111 */
    double distance_from_minimum = v_norm(sep, D) - m->equilibrium_separation;
    return m->prefactor * pow(distance_from_minimum, m->power);
}
static double mh_invert_inside(const mexican_hat *m, double potential) {
    if (m->kind == ECMC_POT_LENNARD_JONES) {
        /* lennard_jones_potential.py:100-115 */
        double sigma_over_r_six = (1 + pow(1 + 4 * potential / m->prefactor, 0.5)) / 2;
        return m->characteristic_length / pow(sigma_over_r_six, 1.0 / 6.0);
    }
    /* displaced_even_power_potential.py:113-127 */
    return m->equilibrium_separation - pow(potential / m->prefactor, m->inverse_power_);
}
static double mh_invert_outside(const mexican_hat *m, double potential) {
    if (m->kind == ECMC_POT_LENNARD_JONES) {
        /* lennard_jones_potential.py:117-135 */
        if (potential >= 0.0) return ORC_INF;
        double sigma_over_r_six = (1 - pow(1 + 4 * potential / m->prefactor, 0.5)) / 2;
        return m->characteristic_length / pow(sigma_over_r_six, 1.0 / 6.0);
    }
    /* displaced_even_power_potential.py:129-142 */
    return m->equilibrium_separation + pow(potential / m->prefactor, m->inverse_power_);
}
static double mh_derivative(const mexican_hat *m, int dir, const double *sep, int D) {
    if (m->kind == ECMC_POT_LENNARD_JONES) {
        /* lennard_jones_potential.py:63-81 */
        return ip_derivative(&m->six, dir, sep, D, 1.0, 1.0) + ip_derivative(&m->twelve, dir, sep, D, 1.0, 1.0);
    }
    /* displaced_even_power_potential.py:73-94 */
    double n = v_norm(sep, D);
    return -m->power * m->prefactor * pow(n - m->equilibrium_separation, m->power - 1) * sep[dir] / n;
}

static double mh_front_inside(const mexican_hat *m, int dir, double dU, double *sep, int D, int *err);

/* _displacement_front_outside_sphere, abstracts.py:384-410 */
static double mh_front_outside(const mexican_hat *m, int dir, double current_potential, double dU, const double *sep,
                               int D, int *err) {
    double new_norm = mh_invert_outside(m, current_potential + dU);
    return disp_new_norm_sq_negative(sep, D, new_norm * new_norm, dir, err);
}
/* _displacement_behind_inside_sphere, abstracts.py:489-530 */
static double mh_behind_inside(const mexican_hat *m, int dir, double current_potential, double dU, double *sep, int D,
                               int *err) {
    double at_max[ECMC_MAX_DIM];
    memcpy(at_max, sep, sizeof(double) * D);
    at_max[dir] = 0.0;
    double maximum_potential_inside = mh_potential(m, at_max, D);
    double potential_difference = maximum_potential_inside - current_potential;
    double displacement;
    if (dU < potential_difference) {
        double new_norm = mh_invert_inside(m, current_potential + dU);
        displacement = disp_new_norm_sq_positive(sep, D, new_norm * new_norm, dir, err);
    } else {
        displacement = sep[dir];
        sep[dir] = 0.0;
        dU -= potential_difference;
        displacement += mh_front_inside(m, dir, dU, sep, D, err);
    }
    return displacement;
}
/* _displacement_front_inside_sphere, abstracts.py:456-487 */
static double mh_front_inside(const mexican_hat *m, int dir, double dU, double *sep, int D, int *err) {
    double displacement = disp_new_norm_sq_negative(sep, D, m->equilibrium_separation_squared, dir, err);
    if (*err) return NAN;
    sep[dir] -= displacement;
    double current_potential = mh_potential(m, sep, D);
    displacement += mh_front_outside(m, dir, current_potential, dU, sep, D, err);
    return displacement;
}
/* _displacement_behind_outside_sphere, abstracts.py:412-454: the try block covers every call, the
 * except ValueError branch restarts from the (possibly already modified) separation. */
static double mh_behind_outside(const mexican_hat *m, int dir, double dU, double *sep, int D, int *err) {
    int local_err = 0;
    double displacement = disp_new_norm_sq_positive(sep, D, m->equilibrium_separation_squared, dir, &local_err);
    if (!local_err) {
        sep[dir] -= displacement;
        double current_potential = mh_potential(m, sep, D);
        displacement += mh_behind_inside(m, dir, current_potential, dU, sep, D, &local_err);
    }
    if (local_err) {
        displacement = sep[dir];
        sep[dir] = 0.0;
        double current_potential = mh_potential(m, sep, D);
        displacement += mh_front_outside(m, dir, current_potential, dU, sep, D, err);
    }
    return displacement;
}
/* standard_velocity_displacement, abstracts.py:336-382 */
static double mh_displacement(const mexican_hat *m, int dir, double *sep, int D, double dU, int *err) {
    double norm_of_separation = v_norm(sep, D);
    double displacement;
    if (norm_of_separation >= m->equilibrium_separation) {
        if (sep[dir] <= 0.0) {
            double current_potential = mh_potential(m, sep, D);
            displacement = mh_front_outside(m, dir, current_potential, dU, sep, D, err);
        } else {
            displacement = mh_behind_outside(m, dir, dU, sep, D, err);
        }
    } else {
        if (sep[dir] <= 0.0) {
            displacement = mh_front_inside(m, dir, dU, sep, D, err);
        } else {
            double current_potential = mh_potential(m, sep, D);
            displacement = mh_behind_inside(m, dir, current_potential, dU, sep, D, err);
        }
    }
    return displacement;
}

/* ================================================================================================== */
/* potential/hard_sphere_potential.py, hard_dipole_potential.py (general velocity)                    */
/* ================================================================================================== */
static double v_dot(const double *a, const double *b, int D) {
    pysum s;
    pysum_init(&s);
    for (int d = 0; d < D; d++) pysum_add(&s, a[d] * b[d]);
    return pysum_result(&s);
}
/* HardSpherePotential.displacement, hard_sphere_potential.py:65-99 */
static double hs_displacement(double radius, const double *velocity, const double *sep, int D) {
    double diameter_squared = 4.0 * radius * radius;
    double velocity_squared = v_norm_sq(velocity, D);
    double separation_squared = v_norm_sq(sep, D);
    double vds = v_dot(velocity, sep, D);
    double square_root_term = vds * vds - velocity_squared * (separation_squared - diameter_squared);
    return (square_root_term >= 0.0 && vds >= 0.0) ? (vds - sqrt(square_root_term)) / velocity_squared : ORC_INF;
}
/* HardDipolePotential.displacement, hard_dipole_potential.py:75-114 */
static double hd_displacement(double min_sep, double max_sep, const double *velocity, const double *sep, int D) {
    double min_sq = min_sep * min_sep, max_sq = max_sep * max_sep;
    double velocity_squared = v_norm_sq(velocity, D);
    double separation_squared = v_norm_sq(sep, D);
    double vds = v_dot(velocity, sep, D);
    if (vds >= 0.0) {
        double minimum_term = vds * vds - velocity_squared * (separation_squared - min_sq);
        if (minimum_term >= 0.0) return (vds - sqrt(minimum_term)) / velocity_squared;
    }
    double maximum_term = vds * vds - velocity_squared * (separation_squared - max_sq);
    return (vds + sqrt(maximum_term)) / velocity_squared;
}

/* ================================================================================================== */
/* potential/merged_image_coulomb_potential/merged_image_coulomb_potential.c                          */
/* ================================================================================================== */
typedef struct {
    int fourier_cutoff, fourier_cutoff_sq, position_cutoff, position_cutoff_sq;
    double alpha_over_length, alpha_over_length_sq, two_alpha_over_length_root_pi, system_length,
        two_pi_over_length;
    double prefactor;
    double *fourier_array; /* [(fc+1)^3], index (i*(fc+1)+j)*(fc+1)+k */
} mic_potential;

/* construct_merged_image_coulomb_potential, merged_image_coulomb_potential.c:77-119 */
static int mic_make(mic_potential *p, double prefactor, double alpha, int fourier_cutoff, int position_cutoff,
                    double system_length) {
    int n = fourier_cutoff + 1;
    p->fourier_array = (double *)calloc((size_t)n * n * n, sizeof(double));
    if (!p->fourier_array) return -1;
    for (int k = 0; k < n; k++)
        for (int j = 0; j < n; j++)
            for (int i = 1; i < n; i++) {
                double coefficient;
                if (j == 0 && k == 0) coefficient = 1.0;
                else if (k == 0 || j == 0) coefficient = 2.0;
                else coefficient = 4.0;
                double norm_sq = i * i + j * j + k * k;
                p->fourier_array[(i * n + j) * n + k] =
                    4.0 * i * coefficient / (norm_sq * system_length * system_length)
                    * exp(-M_PI * M_PI * norm_sq / (alpha * alpha));
            }
    p->fourier_cutoff = fourier_cutoff;
    p->fourier_cutoff_sq = fourier_cutoff * fourier_cutoff;
    p->position_cutoff = position_cutoff;
    p->position_cutoff_sq = position_cutoff * position_cutoff;
    p->alpha_over_length = alpha / system_length;
    p->alpha_over_length_sq = alpha * alpha / (system_length * system_length);
    p->two_alpha_over_length_root_pi = 2.0 * alpha / (system_length * sqrt(M_PI));
    p->system_length = system_length;
    p->two_pi_over_length = 2.0 * M_PI / system_length;
    p->prefactor = prefactor;
    return 0;
}
/* derivative, merged_image_coulomb_potential.c:205-274 */
static double mic_c_derivative(const mic_potential *p, double sx, double sy, double sz) {
    double derivative = 0.0;
    double vector_norm, vector_sq, vector_x, vector_y_sq, vector_z_sq;
    int cutoff_x, cutoff_y;
    int i, j, k;
    int n = p->fourier_cutoff + 1;
    for (k = -p->position_cutoff; k < p->position_cutoff + 1; k++) {
        vector_z_sq = (sz + k * p->system_length) * (sz + k * p->system_length);
        cutoff_y = (int)sqrt(p->position_cutoff_sq - k * k);
        for (j = -cutoff_y; j < cutoff_y + 1; j++) {
            vector_y_sq = (sy + j * p->system_length) * (sy + j * p->system_length);
            cutoff_x = (int)sqrt(p->position_cutoff_sq - j * j - k * k);
            for (i = -cutoff_x; i < cutoff_x + 1; i++) {
                vector_x = sx + i * p->system_length;
                vector_sq = vector_x * vector_x + vector_y_sq + vector_z_sq;
                vector_norm = sqrt(vector_sq);
                derivative += vector_x * (p->two_alpha_over_length_root_pi
                                          * exp(-p->alpha_over_length_sq * vector_sq)
                                          + erfc(p->alpha_over_length * vector_norm) / vector_norm) / vector_sq;
            }
        }
    }
    double delta_cos_x = cos(p->two_pi_over_length * sx);
    double delta_sin_x = sin(p->two_pi_over_length * sx);
    double delta_cos_y = cos(p->two_pi_over_length * sy);
    double delta_sin_y = sin(p->two_pi_over_length * sy);
    double delta_cos_z = cos(p->two_pi_over_length * sz);
    double delta_sin_z = sin(p->two_pi_over_length * sz);
    double cos_x = delta_cos_x, sin_x = delta_sin_x;
    double cos_y = 1.0, sin_y = 0.0, cos_z = 1.0, sin_z = 0.0;
    double store_cos_value;
    for (i = 1; i < p->fourier_cutoff + 1; i++) {
        cutoff_y = (int)sqrt(p->fourier_cutoff_sq - i * i);
        for (j = 0; j < cutoff_y + 1; j++) {
            cutoff_x = (int)sqrt(p->fourier_cutoff_sq - i * i - j * j);
            for (k = 0; k < cutoff_x + 1; k++) {
                derivative += p->fourier_array[(i * n + j) * n + k] * sin_x * cos_y * cos_z;
                if (k != cutoff_x) {
                    store_cos_value = cos_z;
                    cos_z = store_cos_value * delta_cos_z - sin_z * delta_sin_z;
                    sin_z = sin_z * delta_cos_z + store_cos_value * delta_sin_z;
                } else if (j != cutoff_y) {
                    store_cos_value = cos_y;
                    cos_y = store_cos_value * delta_cos_y - sin_y * delta_sin_y;
                    sin_y = sin_y * delta_cos_y + store_cos_value * delta_sin_y;
                    cos_z = 1.0;
                    sin_z = 0.0;
                } else if (i != p->fourier_cutoff) {
                    store_cos_value = cos_x;
                    cos_x = store_cos_value * delta_cos_x - sin_x * delta_sin_x;
                    sin_x = sin_x * delta_cos_x + store_cos_value * delta_sin_x;
                    cos_y = 1.0;
                    sin_y = 0.0;
                    cos_z = 1.0;
                    sin_z = 0.0;
                }
            }
        }
    }
    return derivative;
}
/* permutation_3d, base/vectors.py:217-239 */
static void permutation_3d(const double *v, int dir, double *out) {
    out[0] = v[dir % 3]; out[1] = v[(dir + 1) % 3]; out[2] = v[(dir + 2) % 3];
}
/* MergedImageCoulombPotential.standard_velocity_derivative, merged_image_coulomb_potential.py:128-154 */
static double mic_derivative(const mic_potential *p, int dir, const double *sep, double c1, double c2) {
    double s[3];
    permutation_3d(sep, dir, s);
    return p->prefactor * c1 * c2 * mic_c_derivative(p, s[0], s[1], s[2]);
}

/* ================================================================================================== */
/* potential/inverse_power_coulomb_bounding_potential/inverse_power_coulomb_bounding_potential.c      */
/* ================================================================================================== */
/* derivative, .c:53-55 */
static double ipcb_c_derivative(double pp, double sx, double sy, double sz) {
    return pp * sx / pow(sx * sx + sy * sy + sz * sz, 3.0 / 2.0);
}
/* potential, .c:66-68 */
static double ipcb_c_potential(double pp, double sx, double sy, double sz) {
    return pp / sqrt(sx * sx + sy * sy + sz * sz);
}
/* displacement, .c:85-139 */
static double ipcb_c_displacement(double pp, double sx, double sy, double sz, double potential_change, double L) {
    double half = L / 2.0;
    double current_potential = ipcb_c_potential(pp, sx, sy, sz);
    double potential_zero = ipcb_c_potential(pp, 0.0, sy, sz);
    double potential_half_length = ipcb_c_potential(pp, half, sy, sz);
    double per_length = fabs(potential_zero - potential_half_length);
    double displacement = floor(potential_change / per_length) * L;
    potential_change = fmod(potential_change, per_length);
    double new_norm;
    if (pp > 0.0) {
        if (sx <= 0.0) {
            displacement += half + sx;
            sx = half;
            current_potential = potential_half_length;
        } else {
            if (potential_change >= potential_zero - current_potential) {
                potential_change -= (potential_zero - current_potential);
                displacement += sx + half;
                sx = half;
                current_potential = potential_half_length;
            }
        }
        new_norm = pp / (current_potential + potential_change);
        displacement += (sx - sqrt(new_norm * new_norm - (sy * sy + sz * sz)));
    } else {
        if (sx > 0.0) {
            displacement += sx;
            sx = 0.0;
            current_potential = potential_zero;
        } else {
            if (potential_change >= potential_half_length - current_potential) {
                potential_change -= (potential_half_length - current_potential);
                displacement += sx + L;
                sx = 0.0;
                current_potential = potential_zero;
            }
        }
        new_norm = pp / (current_potential + potential_change);
        displacement += (sx + sqrt(new_norm * new_norm - (sy * sy + sz * sz)));
    }
    return displacement;
}
/* InversePowerCoulombBoundingPotential wrappers, inverse_power_coulomb_bounding_potential.py:84-140 */
static double ipcb_derivative(double prefactor, int dir, const double *sep, double c1, double c2) {
    double s[3];
    permutation_3d(sep, dir, s);
    return ipcb_c_derivative(prefactor * c1 * c2, s[0], s[1], s[2]);
}
static double ipcb_displacement(double prefactor, int dir, const double *sep, double c1, double c2, double dU,
                                double L) {
    double s[3];
    permutation_3d(sep, dir, s);
    return ipcb_c_displacement(prefactor * c1 * c2, s[0], s[1], s[2], dU, L);
}

/* ================================================================================================== */
/* Generic potential object driven by EcmcPotential                                                   */
/* ================================================================================================== */
typedef struct {
    int kind;
    inverse_power ip;
    mexican_hat mh;
    mic_potential mic;
    double p0, p1; /* hard sphere radius / dipole min,max / ipcb prefactor */
    double L;
} opotential;

static int pot_make(opotential *o, const EcmcPotential *p, double L) {
    memset(o, 0, sizeof(*o));
    o->kind = p->kind;
    o->L = L;
    switch (p->kind) {
    case ECMC_POT_NONE: return 0;
    case ECMC_POT_INVERSE_POWER: o->ip = ip_make(p->params[0], p->params[1]); return 0;
    case ECMC_POT_LENNARD_JONES: o->mh = lj_make(p->params[0], p->params[1]); return 0;
    case ECMC_POT_DISPLACED_EVEN_POWER: o->mh = dep_make(p->params[0], p->params[1], p->params[2]); return 0;
    case ECMC_POT_HARD_SPHERE: o->p0 = p->params[0]; return 0;
    case ECMC_POT_HARD_DIPOLE: o->p0 = p->params[0]; o->p1 = p->params[1]; return 0;
    case ECMC_POT_MERGED_IMAGE_COULOMB:
        return mic_make(&o->mic, p->params[0], p->params[1], (int)p->params[2], (int)p->params[3], L);
    case ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING: o->p0 = p->params[0]; return 0;
    case ECMC_POT_BENDING: o->p0 = p->params[0]; o->p1 = p->params[1]; return 0;
    default: return -1;
    }
}
static void pot_free(opotential *o) {
    if (o->kind == ECMC_POT_MERGED_IMAGE_COULOMB) free(o->mic.fourier_array);
    o->mic.fourier_array = NULL;
}
/* Potential.derivative for a standard velocity (StandardVelocityPotential.derivative, abstracts.py:80-103):
 * standard_velocity_derivative(direction, ...) * speed */
static double pot_derivative(const opotential *o, int dir, double speed, const double *sep, int D, double c1,
                             double c2) {
    switch (o->kind) {
    case ECMC_POT_INVERSE_POWER: return ip_derivative(&o->ip, dir, sep, D, c1, c2) * speed;
    case ECMC_POT_LENNARD_JONES:
    case ECMC_POT_DISPLACED_EVEN_POWER: return mh_derivative(&o->mh, dir, sep, D) * speed;
    case ECMC_POT_MERGED_IMAGE_COULOMB: return mic_derivative(&o->mic, dir, sep, c1, c2) * speed;
    case ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING: return ipcb_derivative(o->p0, dir, sep, c1, c2) * speed;
    default: return NAN;
    }
}
/* BendingPotential.derivative (bending_potential.py:60-138): time derivatives with respect to the units i, j, k for
 * separation_one = r_i - r_j, separation_two = r_k - r_j. vectors.norm = sqrt(sum of squares), sum() compensated. */
static void bending_derivative(double prefactor, double equilibrium_angle, int dir, double speed, const double *s1,
                               const double *s2, int D, double out[3]) {
    double n1 = v_norm(s1, D), n2 = v_norm(s2, D);
    pysum dot;
    pysum_init(&dot);
    for (int i = 0; i < D; i++) pysum_add(&dot, s1[i] * s2[i]);
    double cosine = pysum_result(&dot) / n1 / n2;
    double angle = acos(cosine);
    double du_dangle = prefactor * (angle - equilibrium_angle);
    double dangle_dcos = -1.0 / sin(angle);
    double dcos_ds1 = s2[dir] / n1 / n2 - cosine * s1[dir] / pow(n1, 2.0);
    double dcos_ds2 = s1[dir] / n1 / n2 - cosine * s2[dir] / pow(n2, 2.0);
    double du_ds1 = du_dangle * dangle_dcos * dcos_ds1;
    double du_ds2 = du_dangle * dangle_dcos * dcos_ds2;
    out[0] = du_ds1 * speed;
    out[1] = (-du_ds1 - du_ds2) * speed;
    out[2] = du_ds2 * speed;
}

/* InvertiblePotential.displacement: time displacement. For standard-velocity potentials
 * standard_velocity_displacement(...) / speed (abstracts.py:212-243); hard potentials take the velocity. */
static double pot_displacement(const opotential *o, int dir, double speed, double *sep, int D, double c1, double c2,
                               double dU) {
    int err = 0;
    double velocity[ECMC_MAX_DIM] = {0.0, 0.0, 0.0};
    velocity[dir] = speed;
    switch (o->kind) {
    case ECMC_POT_INVERSE_POWER: return ip_displacement(&o->ip, dir, sep, D, c1, c2, dU, &err) / speed;
    case ECMC_POT_LENNARD_JONES:
    case ECMC_POT_DISPLACED_EVEN_POWER: return mh_displacement(&o->mh, dir, sep, D, dU, &err) / speed;
    case ECMC_POT_HARD_SPHERE: return hs_displacement(o->p0, velocity, sep, D);
    case ECMC_POT_HARD_DIPOLE: return hd_displacement(o->p0, o->p1, velocity, sep, D);
    case ECMC_POT_INVERSE_POWER_COULOMB_BOUNDING: return ipcb_displacement(o->p0, dir, sep, c1, c2, dU, o->L) / speed;
    default: return NAN;
    }
}
static int pot_needs_potential_change(int kind) {
    return !(kind == ECMC_POT_HARD_SPHERE || kind == ECMC_POT_HARD_DIPOLE);
}

/* StandardVelocityPotential._analyse_velocity, abstracts.py:105-140: exactly one non-zero component, which is
 * positive. Returns the direction or -1. */
static int analyse_velocity(const double *velocity, int D, double *speed) {
    int dir = -1;
    for (int d = 0; d < D; d++)
        if (velocity[d] != 0.0) {
            if (dir >= 0) return -1;
            dir = d;
        }
    if (dir < 0 || !(velocity[dir] > 0.0)) return -1;
    *speed = velocity[dir];
    return dir;
}
static int pot_is_hard(int kind) { return kind == ECMC_POT_HARD_SPHERE || kind == ECMC_POT_HARD_DIPOLE; }

/* Entry points with the reference's signatures: derivative(velocity, separation, charges) and
 * displacement(velocity, separation, charges, potential_change) (potential.py:154-301). */
ORC_API void orc_potential_derivative_batch(const EcmcPotential *p, int dimension, double L, const double *velocity,
                                            size_t n, const double *seps, const double *charges, double *out) {
    opotential o;
    double speed = 0.0;
    int dir = analyse_velocity(velocity, dimension, &speed);
    if (pot_make(&o, p, L) || dir < 0) {
        for (size_t i = 0; i < n; i++) out[i] = NAN;
        return;
    }
    for (size_t i = 0; i < n; i++) {
        double c1 = charges ? charges[2 * i] : 1.0, c2 = charges ? charges[2 * i + 1] : 1.0;
        out[i] = pot_derivative(&o, dir, speed, seps + i * dimension, dimension, c1, c2);
    }
    pot_free(&o);
}
ORC_API void orc_potential_displacement_batch(const EcmcPotential *p, int dimension, double L, const double *velocity,
                                              size_t n, const double *seps, const double *charges, const double *dUs,
                                              double *out) {
    opotential o;
    double speed = 0.0;
    int dir = analyse_velocity(velocity, dimension, &speed);
    if (pot_make(&o, p, L) || (dir < 0 && !pot_is_hard(p->kind))) {
        for (size_t i = 0; i < n; i++) out[i] = NAN;
        return;
    }
    for (size_t i = 0; i < n; i++) {
        double s[ECMC_MAX_DIM] = {0, 0, 0};
        memcpy(s, seps + i * dimension, sizeof(double) * dimension);
        double c1 = charges ? charges[2 * i] : 1.0, c2 = charges ? charges[2 * i + 1] : 1.0;
        if (p->kind == ECMC_POT_HARD_SPHERE) out[i] = hs_displacement(o.p0, velocity, s, dimension);
        else if (p->kind == ECMC_POT_HARD_DIPOLE) out[i] = hd_displacement(o.p0, o.p1, velocity, s, dimension);
        else out[i] = pot_displacement(&o, dir, speed, s, dimension, c1, c2, dUs ? dUs[i] : 0.0);
    }
    pot_free(&o);
}

/* ================================================================================================== */
/* Random stream (spec in DESIGN.md): Philox4x32-10, key = (stream, seed),                            */
/* counter = (event_lo, event_hi, slot, block). Doubles are built like CPython's random():            */
/* (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53.                                                         */
/* ================================================================================================== */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
static uint32_t rng_word(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t index) {
    uint32_t c[4] = {(uint32_t)event, (uint32_t)(event >> 32), slot, index >> 2};
    philox4x32_10(c, stream, seed);
    return c[index & 3];
}
static double rng_double(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t index) {
    uint32_t c[4] = {(uint32_t)event, (uint32_t)(event >> 32), slot, index >> 1};
    philox4x32_10(c, stream, seed);
    uint32_t a = c[2 * (index & 1)] >> 5, b = c[2 * (index & 1) + 1] >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}
ORC_API void orc_random_doubles(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t first,
                                size_t n, double *out) {
    for (size_t i = 0; i < n; i++) out[i] = rng_double(seed, stream, event, slot, first + (uint32_t)i);
}
ORC_API void orc_random_words(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t first,
                              size_t n, uint32_t *out) {
    for (size_t i = 0; i < n; i++) out[i] = rng_word(seed, stream, event, slot, first + (uint32_t)i);
}
/* random.expovariate(lambd) = -log(1 - random()) / lambd (CPython Lib/random.py) */
static double rng_expovariate(double u, double lambd) { return -log(1.0 - u) / lambd; }
/* random._randbelow_with_getrandbits(n): k = n.bit_length(); r = getrandbits(k) until r < n.
 * getrandbits(k) of this stream = next word >> (32 - k). */
static uint32_t rng_randbelow(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t n) {
    int k = 0;
    while ((n >> k) != 0) k++;
    uint32_t index = 0;
    for (;;) {
        uint32_t r = rng_word(seed, stream, event, slot, index++) >> (32 - k);
        if (r < n) return r;
    }
}

/* the same, continuing in the word stream at *index (successive randint calls of one handler) */
static uint32_t rng_randbelow_from(uint32_t seed, uint32_t stream, uint64_t event, uint32_t slot, uint32_t n,
                                   uint32_t *index) {
    int k = 0;
    while ((n >> k) != 0) k++;
    for (;;) {
        uint32_t r = rng_word(seed, stream, event, slot, (*index)++) >> (32 - k);
        if (r < n) return r;
    }
}

/* ================================================================================================== */
/* activator/internal_state/cell_occupancy/cells: CuboidPeriodicCells                                 */
/* ================================================================================================== */
typedef struct {
    int D;
    int per_side[ECMC_MAX_DIM];
    int cumulative[ECMC_MAX_DIM];
    int n_cells;
    int neighbor_layers;
    double side_length[ECMC_MAX_DIM];
    double L;
    double *cell_min; /* [n_cells][D] */
    double *cell_max;
} ocells;

static double next_float_up(double x) {
    /* cuboid_cells.py:_next_float_up */
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (x == 0.0) x = 0.0;
    int64_t n;
    memcpy(&n, &x, 8);
    if (n >= 0) n += 1; else n -= 1;
    memcpy(&x, &n, 8);
    return x;
}
static double next_float_down(double x) { return -next_float_up(-x); }

static void cell_identifier(const ocells *c, int cell, int *id) {
    for (int d = 0; d < c->D; d++) id[d] = (cell / c->cumulative[d]) % c->per_side[d];
}
/* CuboidCells.__init__, cuboid_cells.py:95-146 */
static int cells_make(ocells *c, int D, const int *per_side, int neighbor_layers, double L) {
    c->D = D; c->L = L; c->neighbor_layers = neighbor_layers;
    c->n_cells = 1;
    for (int d = 0; d < D; d++) {
        c->per_side[d] = per_side[d];
        c->side_length[d] = L / per_side[d];
        c->cumulative[d] = c->n_cells;
        c->n_cells *= per_side[d];
    }
    c->cell_min = (double *)malloc(sizeof(double) * c->n_cells * D);
    c->cell_max = (double *)malloc(sizeof(double) * c->n_cells * D);
    if (!c->cell_min || !c->cell_max) return -1;
    for (int cell = 0; cell < c->n_cells; cell++) {
        int id[ECMC_MAX_DIM];
        cell_identifier(c, cell, id);
        for (int d = 0; d < D; d++) {
            double h = c->side_length[d];
            double lower = id[d] * h;
            double upper = (id[d] + 1) * h;
            if (lower > 0.0) {
                while ((int)(lower / h) == id[d]) lower = next_float_down(lower);
                while ((int)(lower / h) < id[d]) lower = next_float_up(lower);
            }
            while ((int)(upper / h) == id[d]) upper = next_float_up(upper);
            while ((int)(upper / h) > id[d]) upper = next_float_down(upper);
            c->cell_min[cell * D + d] = lower;
            c->cell_max[cell * D + d] = upper;
        }
    }
    return 0;
}
static void cells_free(ocells *c) { free(c->cell_min); free(c->cell_max); c->cell_min = c->cell_max = NULL; }
/* position_to_cell, cuboid_cells.py:190-211: int(x / cell_len) truncates */
static int position_to_cell(const ocells *c, const double *pos) {
    int cell = 0;
    for (int d = 0; d < c->D; d++) cell += (int)(pos[d] / c->side_length[d]) * c->cumulative[d];
    return cell;
}
/* CuboidPeriodicCells.neighbor_cell (positive direction), cuboid_periodic_cells.py:102-139 */
static int neighbor_cell_positive(const ocells *c, int cell, int dir) {
    int id[ECMC_MAX_DIM];
    cell_identifier(c, cell, id);
    int n = 0;
    for (int d = 0; d < c->D; d++)
        n += (d != dir ? id[d] : (id[d] + 1) % c->per_side[d]) * c->cumulative[d];
    return n;
}
/* translate, cuboid_periodic_cells.py:182-207 */
static int cells_translate(const ocells *c, int cell, int relative_cell) {
    double p[ECMC_MAX_DIM];
    for (int d = 0; d < c->D; d++)
        p[d] = correct_position_entry((c->cell_max[cell * c->D + d] + c->cell_min[cell * c->D + d]) / 2.0
                                      + c->cell_min[relative_cell * c->D + d], c->L);
    return position_to_cell(c, p);
}
/* relative_cell, cuboid_periodic_cells.py:155-180 */
static int cells_relative(const ocells *c, int cell, int reference_cell) {
    double p[ECMC_MAX_DIM];
    for (int d = 0; d < c->D; d++)
        p[d] = correct_position_entry((c->cell_max[cell * c->D + d] + c->cell_min[cell * c->D + d]) / 2.0
                                      - c->cell_min[reference_cell * c->D + d], c->L);
    return position_to_cell(c, p);
}
/* _yield_nearby_cells, cuboid_periodic_cells.py:74-100; returns count, duplicates removed (it is a set) */
static int nearby_cells(const ocells *c, int cell, int *out) {
    int id[ECMC_MAX_DIM], off[ECMC_MAX_DIM];
    int nl = c->neighbor_layers, w = 2 * nl + 1, total = 1, count = 0;
    cell_identifier(c, cell, id);
    for (int d = 0; d < c->D; d++) total *= w;
    for (int t = 0; t < total; t++) {
        int rem = t, idx = 0;
        for (int d = 0; d < c->D; d++) { off[d] = rem % w - nl; rem /= w; }
        for (int d = 0; d < c->D; d++) {
            int v = (id[d] + off[d]) % c->per_side[d];
            if (v < 0) v += c->per_side[d];
            idx += v * c->cumulative[d];
        }
        int dup = 0;
        for (int i = 0; i < count; i++) if (out[i] == idx) { dup = 1; break; }
        if (!dup) out[count++] = idx;
    }
    return count;
}

ORC_API int orc_cells_geometry(int D, const int *per_side, double L, double *cell_min, double *cell_max) {
    ocells c;
    if (cells_make(&c, D, per_side, 1, L)) return -1;
    memcpy(cell_min, c.cell_min, sizeof(double) * c.n_cells * D);
    memcpy(cell_max, c.cell_max, sizeof(double) * c.n_cells * D);
    cells_free(&c);
    return 0;
}
ORC_API int orc_position_to_cell(int D, const int *per_side, double L, const double *pos) {
    ocells c;
    c.D = D; c.L = L;
    int n = 1;
    for (int d = 0; d < D; d++) { c.per_side[d] = per_side[d]; c.side_length[d] = L / per_side[d]; c.cumulative[d] = n; n *= per_side[d]; }
    return position_to_cell(&c, pos);
}
ORC_API int orc_cells_translate(int D, const int *per_side, double L, int cell, int relative_cell) {
    ocells c;
    if (cells_make(&c, D, per_side, 1, L)) return -1;
    int r = cells_translate(&c, cell, relative_cell);
    cells_free(&c);
    return r;
}
ORC_API int orc_cells_relative(int D, const int *per_side, double L, int cell, int reference_cell) {
    ocells c;
    if (cells_make(&c, D, per_side, 1, L)) return -1;
    int r = cells_relative(&c, cell, reference_cell);
    cells_free(&c);
    return r;
}
ORC_API int orc_nearby_cells(int D, const int *per_side, int neighbor_layers, double L, int cell, int *out) {
    ocells c;
    c.D = D; c.L = L; c.neighbor_layers = neighbor_layers;
    int n = 1;
    for (int d = 0; d < D; d++) { c.per_side[d] = per_side[d]; c.cumulative[d] = n; n *= per_side[d]; }
    return nearby_cells(&c, cell, out);
}

/* ================================================================================================== */
/* The chain: state + one iteration of the mediator loop                                              */
/* ================================================================================================== */
typedef struct OrcChain {
    EcmcProgram prog;
    int D, N;
    double L;
    ocells cells;
    opotential pair_pot, pair_bound, veto_pot, bond_pot, inter_pot, bending_pot;
    int molecules; /* composite objects in root-level cells (cell_level = 1 with two node levels) */
    int npr;          /* nodes per root (1: point masses) */
    double *root_pos; /* [N / npr][D] root-unit positions of composite objects */
    /* copied veto tables */
    EcmcWalkerTable upper[ECMC_MAX_DIM], lower[ECMC_MAX_DIM];
    double *bounds;
    int *nearby_offsets; /* nearby cells of cell zero (relative) */
    unsigned char *is_nearby_of_zero;
    /* state */
    double *pos;    /* [N][D] */
    double *charge; /* [N] */
    int *occ;       /* [n_cells][max_occ] */
    int *surplus;   /* [max_surplus] */
    int n_surplus;
    EcmcChainState st;
    int started;
    EcmcStats stats;
} OrcChain;

static int copy_walker(EcmcWalkerTable *dst, const EcmcWalkerTable *src) {
    *dst = *src;
    if (src->n_entries <= 0) { dst->cell_a = dst->cell_b = NULL; dst->rate_a = NULL; return 0; }
    int32_t *a = (int32_t *)malloc(sizeof(int32_t) * src->n_entries);
    int32_t *b = (int32_t *)malloc(sizeof(int32_t) * src->n_entries);
    double *r = (double *)malloc(sizeof(double) * src->n_entries);
    if (!a || !b || !r) return -1;
    memcpy(a, src->cell_a, sizeof(int32_t) * src->n_entries);
    memcpy(b, src->cell_b, sizeof(int32_t) * src->n_entries);
    memcpy(r, src->rate_a, sizeof(double) * src->n_entries);
    dst->cell_a = a; dst->cell_b = b; dst->rate_a = r;
    return 0;
}

ORC_API void orc_chain_destroy(OrcChain *c) {
    if (!c) return;
    for (int d = 0; d < ECMC_MAX_DIM; d++) {
        free((void *)c->upper[d].cell_a); free((void *)c->upper[d].cell_b); free((void *)c->upper[d].rate_a);
        free((void *)c->lower[d].cell_a); free((void *)c->lower[d].cell_b); free((void *)c->lower[d].rate_a);
    }
    free(c->bounds); free(c->nearby_offsets); free(c->is_nearby_of_zero);
    free(c->pos); free(c->charge); free(c->occ); free(c->surplus); free(c->root_pos);
    cells_free(&c->cells);
    pot_free(&c->pair_pot); pot_free(&c->pair_bound); pot_free(&c->veto_pot); pot_free(&c->bond_pot);
    pot_free(&c->inter_pot); pot_free(&c->bending_pot);
    free(c);
}

ORC_API OrcChain *orc_chain_create(const EcmcProgram *prog) {
    if (!prog || prog->abi_version != ECMC_ABI_VERSION) return NULL;
    OrcChain *c = (OrcChain *)calloc(1, sizeof(OrcChain));
    if (!c) return NULL;
    c->prog = *prog;
    c->D = prog->dimension; c->N = prog->n_particles; c->L = prog->system_length;
    if (cells_make(&c->cells, c->D, prog->cells_per_side, prog->neighbor_layers, c->L)) goto fail;
    if (pot_make(&c->pair_pot, &prog->pair_potential, c->L)) goto fail;
    if (pot_make(&c->pair_bound, &prog->pair_bounding_potential, c->L)) goto fail;
    if (pot_make(&c->veto_pot, &prog->veto_potential, c->L)) goto fail;
    if (pot_make(&c->bond_pot, &prog->bond_potential, c->L)) goto fail;
    if (pot_make(&c->inter_pot, &prog->inter_potential, c->L)) goto fail;
    if (pot_make(&c->bending_pot, &prog->bending_potential, c->L)) goto fail;
    c->npr = prog->nodes_per_root > 1 ? prog->nodes_per_root : 1;
    c->molecules = prog->cell_level == 1 && c->npr > 1;
    if ((prog->pair_handler == ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING || prog->bending_enabled ||
         (prog->n_inter_factors > 0 && !prog->eoc_sequential)) && !c->molecules) goto fail;
    /* general velocities: two-dimensional composite point objects without a cell system, hard potentials only */
    if (prog->eoc_sequential && (c->D != 2 || !prog->no_cells || c->npr < 2 || c->molecules ||
                                 prog->pair_handler != ECMC_PAIR_NONE || prog->veto_enabled)) goto fail;
    if (c->molecules && (c->D != 3 || c->npr > 4 || prog->max_occupants != 1)) goto fail;
    /* a cell system for one kind of leaf only: composite-object pairs from the factor type map, one two-leaf factor
     * between those leaves found through the cells */
    if (prog->cell_child && (!c->molecules || prog->no_cells || prog->cell_child < 1 || prog->cell_child > c->npr ||
                             prog->pair_handler != ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING || prog->root_mode ||
                             prog->n_inter_factors != 1 || prog->inter_factors[0][0] != prog->cell_child - 1 ||
                             prog->inter_factors[0][1] != prog->cell_child - 1 || !prog->boundary_keeps_factors ||
                             (prog->veto_enabled != ECMC_FAR_NONE && prog->veto_enabled != ECMC_FAR_CELL_BOUNDING) ||
                             !(prog->inter_bound_max_displacement > 0.0))) goto fail;
    /* root-unit-active mode: dipoles without a cell system, composite-object pair handler */
    if (prog->root_mode && (!c->molecules || c->npr != 2 || !prog->no_cells || prog->bending_enabled ||
                            prog->pair_handler != ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING)) goto fail;
    if (c->N % c->npr || prog->n_bonds < 0 || prog->n_bonds > ECMC_MAX_BONDS) goto fail;
    if (prog->n_bonds > 0 && c->npr == 1) goto fail;
    c->root_pos = (double *)calloc((size_t)(c->N / c->npr) * c->D, sizeof(double));
    if (!c->root_pos) goto fail;
    c->is_nearby_of_zero = (unsigned char *)calloc(c->cells.n_cells, 1);
    c->nearby_offsets = (int *)malloc(sizeof(int) * c->cells.n_cells);
    if (!c->is_nearby_of_zero || !c->nearby_offsets) goto fail;
    {
        int n = nearby_cells(&c->cells, 0, c->nearby_offsets);
        for (int i = 0; i < n; i++) c->is_nearby_of_zero[c->nearby_offsets[i]] = 1;
    }
    if (prog->veto_enabled) {
        if (!prog->veto_tables) goto fail;
        for (int d = 0; d < c->D && prog->veto_enabled == ECMC_FAR_CELL_VETO; d++) {
            if (copy_walker(&c->upper[d], &prog->veto_tables->upper[d])) goto fail;
            if (copy_walker(&c->lower[d], &prog->veto_tables->lower[d])) goto fail;
        }
        size_t nb = (size_t)c->cells.n_cells * c->D * 2;
        c->bounds = (double *)malloc(sizeof(double) * nb);
        if (!c->bounds) goto fail;
        memcpy(c->bounds, prog->veto_tables->bounds, sizeof(double) * nb);
    }
    c->prog.veto_tables = NULL;
    c->pos = (double *)calloc((size_t)c->N * c->D, sizeof(double));
    c->charge = (double *)malloc(sizeof(double) * c->N);
    c->occ = (int *)malloc(sizeof(int) * c->cells.n_cells * prog->max_occupants);
    c->surplus = (int *)malloc(sizeof(int) * (prog->max_surplus > 0 ? prog->max_surplus : 1));
    if (!c->pos || !c->charge || !c->occ || !c->surplus) goto fail;
    for (int i = 0; i < c->N; i++) c->charge[i] = 1.0;
    return c;
fail:
    orc_chain_destroy(c);
    return NULL;
}

ORC_API void orc_chain_set_positions(OrcChain *c, const double *pos, const double *charge) {
    memcpy(c->pos, pos, sizeof(double) * c->N * c->D);
    if (charge) memcpy(c->charge, charge, sizeof(double) * c->N);
}
ORC_API void orc_chain_get_positions(const OrcChain *c, double *pos) {
    memcpy(pos, c->pos, sizeof(double) * c->N * c->D);
}
ORC_API void orc_chain_set_roots(OrcChain *c, const double *roots) {
    memcpy(c->root_pos, roots, sizeof(double) * (c->N / c->npr) * c->D);
}
ORC_API void orc_chain_get_roots(const OrcChain *c, double *roots) {
    memcpy(roots, c->root_pos, sizeof(double) * (c->N / c->npr) * c->D);
}
ORC_API void orc_chain_get_state(const OrcChain *c, EcmcChainState *st) { *st = c->st; }
ORC_API void orc_chain_set_state(OrcChain *c, const EcmcChainState *st) { c->st = *st; c->started = 1; }
ORC_API void orc_chain_get_cells(const OrcChain *c, int32_t *occ, int32_t *surplus, int32_t *n_surplus) {
    memcpy(occ, c->occ, sizeof(int) * c->cells.n_cells * c->prog.max_occupants);
    memcpy(surplus, c->surplus, sizeof(int) * c->n_surplus);
    *n_surplus = c->n_surplus;
}
ORC_API void orc_chain_set_cells(OrcChain *c, const int32_t *occ, const int32_t *surplus, int32_t n_surplus) {
    memcpy(c->occ, occ, sizeof(int) * c->cells.n_cells * c->prog.max_occupants);
    memcpy(c->surplus, surplus, sizeof(int) * n_surplus);
    c->n_surplus = n_surplus;
}
ORC_API void orc_chain_get_stats(const OrcChain *c, EcmcStats *s) { *s = c->stats; }

/* occupancy helpers: SingleActiveCellOccupancy, single_active_cell_occupancy.py */
static int occ_count(const OrcChain *c, int cell) {
    int n = 0;
    for (int s = 0; s < c->prog.max_occupants; s++) if (c->occ[cell * c->prog.max_occupants + s] >= 0) n++;
    return n;
}
static void occ_insert(OrcChain *c, int cell, int id) {
    /* :117-121 and :176-180: append to occupants if below the maximum, else to surplus */
    if (occ_count(c, cell) < c->prog.max_occupants) {
        for (int s = 0; s < c->prog.max_occupants; s++)
            if (c->occ[cell * c->prog.max_occupants + s] < 0) { c->occ[cell * c->prog.max_occupants + s] = id; return; }
    }
    if (c->n_surplus >= c->prog.max_surplus) { c->stats.capacity_errors++; return; }
    c->surplus[c->n_surplus++] = id;
}
static void occ_remove(OrcChain *c, int cell, int id) {
    /* :186-193: remove from the occupants (the surplus is NOT promoted: the guarding expression
     * `not self._surplus.get(cell, True)` is only true for an empty list, which never exists) or, on
     * ValueError, from the surplus */
    int m = c->prog.max_occupants;
    for (int s = 0; s < m; s++)
        if (c->occ[cell * m + s] == id) {
            /* keep list order like list.remove */
            for (int t = s; t + 1 < m; t++) c->occ[cell * m + t] = c->occ[cell * m + t + 1];
            c->occ[cell * m + m - 1] = -1;
            return;
        }
    for (int s = 0; s < c->n_surplus; s++)
        if (c->surplus[s] == id) {
            for (int t = s; t + 1 < c->n_surplus; t++) c->surplus[t] = c->surplus[t + 1];
            c->n_surplus--;
            return;
        }
    c->stats.capacity_errors++;
}

/* EndOfChainEventHandler.send_event_time for the candidate created at the start of a chain
 * (abstracts/end_of_chain_event_handler.py:80-105; single_independent_active_periodic_direction_...:203-237):
 * new_chain_time = (last_committed - current) + chain_time with last_committed == current at creation. */
static void schedule_end_of_chain(OrcChain *c) {
    otime now = {c->st.time_q, c->st.time_r};
    double new_chain_time = time_sub(now, now) + c->prog.chain_time;
    if (c->prog.root_mode) {
        /* the candidate is also re-created by a RootLeafUnitActiveSwitcher event, between two ends of chain:
         * (_last_committed_event_time - current_time_stamp) + chain_time, :203-215 */
        otime last = {c->st.eoc_last_q, c->st.eoc_last_r};
        new_chain_time = time_sub(last, now) + c->prog.chain_time;
    }
    otime t = time_add(now, new_chain_time);
    c->st.eoc_q = t.q; c->st.eoc_r = t.r;
    if (c->prog.root_mode && c->st.mode == 1) {
        /* the root unit was independent active: (randint(0, number_of_root_nodes - 1),), :226-229; recorded as the
         * first leaf of that object */
        c->st.eoc_next_active = (int)rng_randbelow(c->prog.seed, c->st.stream, c->st.event_counter,
                                                   ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)(c->N / c->npr)) * c->npr;
    } else if (c->npr == 1) {
        c->st.eoc_next_active = (int)rng_randbelow(c->prog.seed, c->st.stream, c->st.event_counter,
                                                   ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)c->N);
    } else {
        /* a leaf unit was active: (randint(0, number_of_root_nodes - 1), randint(0, number_of_nodes_per_root_node - 1)),
         * single_independent_active_periodic_direction_end_of_chain_event_handler.py:231-237; the second randint
         * continues in the word stream where the first one stopped */
        uint32_t index = 0;
        uint32_t root = rng_randbelow_from(c->prog.seed, c->st.stream, c->st.event_counter,
                                           ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)(c->N / c->npr), &index);
        uint32_t child = rng_randbelow_from(c->prog.seed, c->st.stream, c->st.event_counter,
                                            ECMC_SLOT(ECMC_SLOT_END_OF_CHAIN, 0), (uint32_t)c->npr, &index);
        c->st.eoc_next_active = (int)(root * (uint32_t)c->npr + child);
    }
}

/* Start of run: SingleActiveCellOccupancy.initialize (:95-121), InitialChainStartOfRunEventHandler
 * (initial_chain_start_of_run_event_handler.py:92-131), then the first internal-state update (:149-203). */
ORC_API void orc_chain_start(OrcChain *c, uint32_t stream) {
    int m = c->prog.max_occupants;
    for (int i = 0; i < c->cells.n_cells * m; i++) c->occ[i] = -1;
    c->n_surplus = 0;
    if (c->molecules && c->prog.cell_child) {
        /* cell_level = 2 with a charge indicator: only the leaves of one kind are stored (:95-121, _is_relevant_unit) */
        for (int r = 0; r < c->N / c->npr; r++) {
            int leaf = r * c->npr + c->prog.cell_child - 1;
            occ_insert(c, position_to_cell(&c->cells, c->pos + leaf * c->D), leaf);
        }
    } else if (c->molecules) {
        /* cell_level = 1: the cells hold the root units (single_active_cell_occupancy.py:95-121) */
        for (int r = 0; r < c->N / c->npr; r++) occ_insert(c, position_to_cell(&c->cells, c->root_pos + r * c->D), r);
    } else {
        for (int i = 0; i < c->N; i++) occ_insert(c, position_to_cell(&c->cells, c->pos + i * c->D), i);
    }
    memset(&c->st, 0, sizeof(c->st));
    c->st.stream = stream;
    c->st.active = c->prog.initial_active;
    c->st.direction = c->prog.initial_direction;
    c->st.time_q = 0.0; c->st.time_r = 0.0;
    c->st.event_counter = 0;
    if (c->molecules && c->prog.cell_child) {
        /* the occupancy has an active cell only while a stored kind of leaf is active (:149-203) */
        c->st.active_cell = 0;
        if (c->st.active % c->npr == c->prog.cell_child - 1) {
            c->st.active_cell = position_to_cell(&c->cells, c->pos + c->st.active * c->D);
            occ_remove(c, c->st.active_cell, c->st.active);
        }
    } else if (c->molecules) {
        c->st.active_cell = position_to_cell(&c->cells, c->root_pos + (c->st.active / c->npr) * c->D);
        occ_remove(c, c->st.active_cell, c->st.active / c->npr);
    } else {
        c->st.active_cell = position_to_cell(&c->cells, c->pos + c->st.active * c->D);
        occ_remove(c, c->st.active_cell, c->st.active);
    }
    schedule_end_of_chain(c);
    c->st.pending_kind = ECMC_EVENT_NONE;
    if (c->prog.root_mode) {
        /* the leaf-to-root switcher is created at the start of the run: time stamp of the root unit + chain length
         * (RootLeafUnitActiveSwitcher.send_event_time, root_leaf_unit_active_switcher.py:102-127) */
        otime zero = {0.0, 0.0};
        otime t = time_add(zero, c->prog.switch_chain_length[0]);
        c->st.switch_q = t.q; c->st.switch_r = t.r;
    }
    if (c->prog.eoc_sequential) {
        /* InitialChainStartOfRunEventHandler (initial_chain_start_of_run_event_handler.py:92-131): the active leaf gets
         * speed along the initial direction, its root unit that velocity times the leaf's weight
         * (_register_velocity_change_leaf_cnode, abstracts.py:165-190) */
        double weight = 1.0 / c->npr;
        for (int d = 0; d < 2; d++) {
            c->st.velocity[d] = d == c->prog.initial_direction ? c->prog.speed : 0.0;
            c->st.root_velocity[d] = c->st.velocity[d] * weight;
        }
    }
    c->started = 1;
    memset(&c->stats, 0, sizeof(c->stats));
}

typedef struct {
    int kind, target, target_cell;
    otime t;
    double rate;
} candidate;

/* One pair candidate: TwoLeafUnitEventHandler.send_event_time (two_leaf_unit_event_handler.py:105-138) or
 * TwoLeafUnitBoundingPotentialEventHandler.send_event_time (two_leaf_unit_bounding_potential_event_handler.py:112-146) */
static otime pair_candidate_time(OrcChain *c, int target) {
    const double *pa = c->pos + c->st.active * c->D, *pt = c->pos + target * c->D;
    double sep[ECMC_MAX_DIM] = {0, 0, 0};
    separation_vector(pa, pt, c->D, c->L, sep);
    double c1 = c->prog.pair_use_charge ? c->charge[c->st.active] : 1.0;
    double c2 = c->prog.pair_use_charge ? c->charge[target] : 1.0;
    const opotential *pot = c->prog.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING ? &c->pair_bound : &c->pair_pot;
    double dU = 0.0;
    if (pot_needs_potential_change(pot->kind)) {
        double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter,
                              ECMC_SLOT(ECMC_SLOT_PAIR_TIME, target), 0);
        dU = rng_expovariate(u, c->prog.beta);
    }
    double dt = pot_displacement(pot, c->st.direction, c->prog.speed, sep, c->D, c1, c2, dU);
    otime now = {c->st.time_q, c->st.time_r};
    return time_add(now, dt);
}

/* CellVetoEventHandler.send_event_time, abstracts/cell_veto_event_handler.py:200-238; Walker.sample_cell,
 * walker.py:105-118; InnerPointEstimator.charge_correction_factor, inner_point_estimator.py:165-192 */
static candidate veto_candidate(OrcChain *c) {
    candidate cand;
    cand.kind = ECMC_EVENT_CELL_VETO; cand.target = -1;
    int dir = c->st.direction;
    double charge_factor = 1.0;
    if (c->prog.veto_use_charge)
        charge_factor = c->charge[c->st.active] * 1.0 / c->prog.veto_target_charge;
    const EcmcWalkerTable *w;
    int rate_index;
    if (charge_factor > 0.0) { w = &c->upper[dir]; rate_index = 0; }
    else { charge_factor *= -1.0; w = &c->lower[dir]; rate_index = 1; }
    double total_rate = w->total_rate * charge_factor;
    uint32_t e = rng_randbelow(c->prog.seed, c->st.stream, c->st.event_counter,
                               ECMC_SLOT(ECMC_SLOT_VETO_CHOICE, 0), (uint32_t)w->n_entries);
    double u0 = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_VETO_TIME, 0), 0);
    double u1 = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_VETO_TIME, 0), 1);
    /* random.uniform(0.0, mean) = 0.0 + (mean - 0.0) * random() */
    int relative_cell = (0.0 + (w->mean_rate - 0.0) * u0 <= w->rate_a[e]) ? w->cell_a[e] : w->cell_b[e];
    cand.rate = c->bounds[(relative_cell * c->D + dir) * 2 + rate_index] * charge_factor;
    /* the active cell is recomputed from the position, cell_veto_event_handler.py:216 */
    int active_cell = c->molecules ? position_to_cell(&c->cells, c->root_pos + (c->st.active / c->npr) * c->D)
                                   : position_to_cell(&c->cells, c->pos + c->st.active * c->D);
    cand.target_cell = cells_translate(&c->cells, active_cell, relative_cell);
    double dt = rng_expovariate(u1, c->prog.beta) / (total_rate * c->prog.speed);
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, dt);
    return cand;
}

/* TwoLeafUnitCellBoundingPotentialEventHandler.send_event_time (two_leaf_unit_cell_bounding_potential_event_handler.py:
 * 137-177) with CellBoundingPotential.standard_velocity_displacement (cell_bounding_potential.py:155-238): the cells of
 * both units are recomputed from their positions, the bound of the relative cell times the charge correction factor
 * (inner_point_estimator.py:165-192) is a constant event rate. target_cell carries the relative cell. */
static candidate cell_bounding_candidate(OrcChain *c, int target) {
    candidate cand;
    cand.kind = ECMC_EVENT_CELL_BOUNDING; cand.target = target;
    int dir = c->st.direction;
    int active_cell = position_to_cell(&c->cells, c->pos + c->st.active * c->D);
    int target_cell = position_to_cell(&c->cells, c->pos + target * c->D);
    int relative_cell = cells_relative(&c->cells, target_cell, active_cell);
    cand.target_cell = relative_cell;
    double c1 = c->prog.veto_use_charge ? c->charge[c->st.active] : 1.0;
    double c2 = c->prog.veto_use_charge ? c->charge[target] : 1.0;
    double charge_product = c->prog.veto_use_charge ? c1 * c2 / c->prog.veto_target_charge : 1.0;
    /* bounds holds (upper, -lower): lower_bound * charge_product for a negative product */
    if (charge_product > 0.0) cand.rate = c->bounds[(relative_cell * c->D + dir) * 2 + 0] * charge_product;
    else cand.rate = -c->bounds[(relative_cell * c->D + dir) * 2 + 1] * charge_product;
    double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_PAIR_TIME, target), 0);
    double dU = rng_expovariate(u, c->prog.beta);
    double displacement = cand.rate > 0 ? dU / cand.rate : ORC_INF;
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, displacement / c->prog.speed);
    return cand;
}

/* A factor-type-map pair factor inside the active leaf's composite object (FactorTypeMapInStateTagger,
 * factor_type_map_in_state_tagger.py:83-107) handled by a TwoLeafUnitEventHandler (two_leaf_unit_event_handler.py:
 * 105-138) with the bond potential. */
static candidate bond_candidate(OrcChain *c, int target) {
    candidate cand;
    cand.kind = ECMC_EVENT_BOND; cand.target = target; cand.target_cell = -1; cand.rate = 0.0;
    const double *pa = c->pos + c->st.active * c->D, *pt = c->pos + target * c->D;
    double sep[ECMC_MAX_DIM] = {0, 0, 0};
    separation_vector(pa, pt, c->D, c->L, sep);
    double dU = 0.0;
    if (pot_needs_potential_change(c->bond_pot.kind)) {
        double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter,
                              ECMC_SLOT(ECMC_SLOT_FACTOR_TIME, target), 0);
        dU = rng_expovariate(u, c->prog.beta);
    }
    double dt = pot_displacement(&c->bond_pot, c->st.direction, c->prog.speed, sep, c->D, 1.0, 1.0, dU);
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, dt);
    return cand;
}

/* CellBoundaryEventHandler.send_event_time, cell_boundary_event_handler.py:122-156 (positive velocity) */
static candidate boundary_candidate(OrcChain *c, double *boundary_out) {
    candidate cand;
    cand.kind = ECMC_EVENT_CELL_BOUNDARY; cand.target = -1; cand.rate = 0.0;
    int dir = c->st.direction;
    const double *pa = c->pos + c->st.active * c->D;
    int cell = position_to_cell(&c->cells, pa);
    int neighbor = neighbor_cell_positive(&c->cells, cell, dir);
    double neighbor_boundary = c->cells.cell_min[neighbor * c->D + dir];
    double separation = neighbor_boundary - pa[dir];
    if (separation < 0.0) separation = separation + c->L; /* next_image, hypercubic_setting.py:191 */
    double time_to_boundary = separation / c->prog.speed;
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, time_to_boundary);
    cand.target_cell = neighbor;
    *boundary_out = neighbor_boundary;
    return cand;
}

/* BasicEventHandler._time_slice_unit for the active particle, abstracts/abstracts.py:82-95 */
static void time_slice_object(OrcChain *c, otime event_time);

static void time_slice_active(OrcChain *c, otime event_time) {
    if (c->prog.root_mode && c->st.mode == 1) { time_slice_object(c, event_time); return; }
    double *pa = c->pos + c->st.active * c->D;
    otime stamp = {c->st.time_q, c->st.time_r};
    double dt = time_sub(event_time, stamp);
    if (c->prog.eoc_sequential) {
        /* general velocities: every component of the leaf and of its root unit, each with its own velocity */
        double *pr = c->root_pos + (c->st.active / c->npr) * c->D;
        for (int d = 0; d < c->D; d++) {
            pa[d] = correct_position_entry(pa[d] + c->st.velocity[d] * dt, c->L);
            pr[d] = correct_position_entry(pr[d] + c->st.root_velocity[d] * dt, c->L);
        }
        c->st.time_q = event_time.q; c->st.time_r = event_time.r;
        return;
    }
    for (int d = 0; d < c->D; d++) {
        double v = d == c->st.direction ? c->prog.speed : 0.0;
        pa[d] = correct_position_entry(pa[d] + v * dt, c->L);
    }
    if (c->npr > 1) {
        /* the root unit of the active leaf moves with velocity * weight, weight = 1 / number of children
         * (_register_velocity_change_leaf_cnode, abstracts.py:165-190); it carries the same time stamp as the leaf and
         * is time-sliced with it (_time_slice_all_units_in_state, :97-101) */
        double *pr = c->root_pos + (c->st.active / c->npr) * c->D;
        double weight = 1.0 / c->npr;
        for (int d = 0; d < c->D; d++) {
            double v = (d == c->st.direction ? c->prog.speed : 0.0) * weight;
            pr[d] = correct_position_entry(pr[d] + v * dt, c->L);
        }
    }
    c->st.time_q = event_time.q; c->st.time_r = event_time.r;
}

/* The internal-state update the activator runs at the top of the next iteration
 * (SingleActiveCellOccupancy.update, single_active_cell_occupancy.py:149-203) */
static void occupancy_update(OrcChain *c, int new_active) {
    if (new_active != c->st.active) {
        occ_insert(c, c->st.active_cell, c->st.active);
        c->st.active = new_active;
        c->st.active_cell = position_to_cell(&c->cells, c->pos + new_active * c->D);
        occ_remove(c, c->st.active_cell, new_active);
    } else {
        c->st.active_cell = position_to_cell(&c->cells, c->pos + new_active * c->D);
    }
}

static int lt_candidate(const candidate *a, const candidate *b) { return time_lt(a->t, b->t); }

/* ================================================================================================== */
/* Molecules: composite objects in root-level cells (water, C4 of SURVEY.md 8d)                        */
/* ================================================================================================== */
/* Lifting.insert + get_active_identifier of the three schemes (lifting/lifting.py:49-91, inside_first_lifting.py:
 * 39-54, outside_first_lifting.py:39-55, ratio_lifting.py:40-56). Draws come from the out-state's slot in call order. */
typedef struct {
    double negative[8];
    int ids[8];
    int n_negative;
    double random_position, sum_positive;
    int active_recorded;
} olifting;
static void lifting_reset(olifting *l) { memset(l, 0, sizeof(*l)); }
/* the uniform draws of a lifting scheme: from the chain's out-state slot, or (c == NULL, tests of the schemes alone)
 * from a caller-supplied list */
static const double *g_lifting_test_uniforms = NULL;
static double lifting_uniform(OrcChain *c, uint32_t *draw) {
    if (!c) return g_lifting_test_uniforms[(*draw)++];
    return rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), (*draw)++);
}
static void lifting_insert(olifting *l, double rate, int id, int is_active, OrcChain *c, uint32_t *draw) {
    if (rate > 0.0) {
        l->sum_positive += rate;
        if (is_active) {
            l->active_recorded = 1;
            double u = lifting_uniform(c, draw);
            l->random_position += 0.0 + (rate - 0.0) * u; /* random.uniform(0.0, lifting_rate) */
        } else if (!l->active_recorded) {
            l->random_position += rate;
        }
    } else {
        l->negative[l->n_negative] = -rate;
        l->ids[l->n_negative++] = id;
    }
}
static int lifting_get(olifting *l, int kind, OrcChain *c, uint32_t *draw) {
    double position = l->random_position;
    if (kind == ECMC_LIFTING_OUTSIDE_FIRST || kind == ECMC_LIFTING_RATIO) {
        pysum total;
        pysum_init(&total);
        for (int i = 0; i < l->n_negative; i++) pysum_add(&total, l->negative[i]);
        double sum_negative = l->n_negative ? pysum_result(&total) : 0.0;
        if (kind == ECMC_LIFTING_OUTSIDE_FIRST) {
            position = sum_negative - l->random_position;
        } else {
            double u = lifting_uniform(c, draw);
            position = 0.0 + (sum_negative - 0.0) * u;
        }
    }
    double summed = 0.0;
    for (int i = 0; i < l->n_negative; i++) {
        summed += l->negative[i];
        if (position <= summed) return l->ids[i];
    }
    return l->n_negative ? l->ids[l->n_negative - 1] : -1;
}

/* The lifting schemes alone (unittests/test_lifting/test_{inside_first,outside_first,ratio}_lifting.py): insert n units
 * (rate, identifier = index, active flag) in order, then ask for the next active identifier. uniforms: the values of
 * random() behind the scheme's random.uniform calls, in call order. Returns the chosen index, -1 if no unit has a
 * negative rate, -2 if the active unit was not recorded (LiftingSchemeError). */
ORC_API int orc_lifting_choose(int kind, int n, const double *rates, int active_index, const double *uniforms) {
    if (n > 8) return -3;
    olifting lift;
    uint32_t draw = 0;
    lifting_reset(&lift);
    g_lifting_test_uniforms = uniforms;
    for (int i = 0; i < n; i++) lifting_insert(&lift, rates[i], i, i == active_index, NULL, &draw);
    if (!lift.active_recorded) return -2;
    return lifting_get(&lift, kind, NULL, &draw);
}
ORC_API void orc_bending_derivative(double prefactor, double equilibrium_angle, int dir, double speed, const double *s1,
                                    const double *s2, int D, double *out) {
    bending_derivative(prefactor, equilibrium_angle, dir, speed, s1, s2, D, out);
}

/* TwoCompositeObjectSummedBoundingPotentialEventHandler.send_event_time
 * (two_composite_object_summed_bounding_potential_event_handler.py:119-156): minimum over the target leaf units (sorted
 * by identifier, abstracts/composite_objects.py:63-83) of the bounding potential's displacement, one expovariate each */
static candidate composite_pair_candidate(OrcChain *c, int target_root) {
    candidate cand;
    cand.kind = ECMC_EVENT_PAIR; cand.target = target_root; cand.target_cell = -1; cand.rate = 0.0;
    const double *pa = c->pos + c->st.active * c->D;
    double best = ORC_INF;
    for (int k = 0; k < c->npr; k++) {
        int t = target_root * c->npr + k;
        double sep[ECMC_MAX_DIM] = {0, 0, 0};
        separation_vector(pa, c->pos + t * c->D, c->D, c->L, sep);
        double c1 = c->prog.pair_use_charge ? c->charge[c->st.active] : 1.0;
        double c2 = c->prog.pair_use_charge ? c->charge[t] : 1.0;
        double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter,
                              ECMC_SLOT(ECMC_SLOT_PAIR_TIME, target_root), (uint32_t)k);
        double dt = pot_displacement(&c->pair_bound, c->st.direction, c->prog.speed, sep, c->D, c1, c2,
                                     rng_expovariate(u, c->prog.beta));
        if (dt < best) best = dt; /* min() keeps the first of equal values */
    }
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, best);
    return cand;
}

/* TwoCompositeObjectCellBoundingPotentialEventHandler.send_event_time
 * (two_composite_object_cell_bounding_potential_event_handler.py:152-196): the composite object `target_root` in a cell that
 * is not nearby the active ROOT's cell; constant event rate = bound of the relative cell times the charge correction
 * factor of the DipoleMonteCarloEstimator, active charge x max |target charges| (dipole_monte_carlo_estimator.py:158-186;
 * CellBoundingPotential._standard_velocity_displacement_with_charges, cell_bounding_potential.py:155-190) */
static candidate composite_cell_bounding_candidate(OrcChain *c, int target_root) {
    candidate cand;
    cand.kind = ECMC_EVENT_CELL_BOUNDING; cand.target = target_root;
    int dir = c->st.direction, npr = c->npr;
    int active_cell = position_to_cell(&c->cells, c->root_pos + (c->st.active / npr) * c->D);
    int target_cell = position_to_cell(&c->cells, c->root_pos + target_root * c->D);
    int relative_cell = cells_relative(&c->cells, target_cell, active_cell);
    cand.target_cell = relative_cell;
    double charge_product = 1.0;
    if (c->prog.veto_use_charge) {
        double largest = 0.0;
        for (int k = 0; k < npr; k++) {
            double a = fabs(c->charge[target_root * npr + k]);
            if (a > largest) largest = a; /* max(abs(charge) for charge in target_charges) */
        }
        charge_product = c->charge[c->st.active] * largest;
    }
    if (charge_product > 0.0) cand.rate = c->bounds[(relative_cell * c->D + dir) * 2 + 0] * charge_product;
    else cand.rate = -c->bounds[(relative_cell * c->D + dir) * 2 + 1] * charge_product;
    double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_PAIR_TIME, target_root), 0);
    double dU = rng_expovariate(u, c->prog.beta);
    double displacement = cand.rate > 0 ? dU / cand.rate : ORC_INF;
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, displacement / c->prog.speed);
    return cand;
}

/* Leaf-to-leaf factors with a bounding potential between the active leaf and leaf k of another object, each handled by a
 * TwoLeafUnitBoundingPotentialEventHandler (two_leaf_unit_bounding_potential_event_handler.py:112-146) fed by non-local
 * factor type map entries ("[0, 2], Coulomb" ... of factor_set_dipoles_atomic.txt: every leaf of one object with every
 * leaf of the other). The draw is keyed like the composite handler's: (pair time, target object), double k. */
static candidate leaf_pair_candidate(OrcChain *c, int target) {
    candidate cand;
    cand.kind = ECMC_EVENT_PAIR; cand.target = target; cand.target_cell = -1; cand.rate = 0.0;
    double sep[ECMC_MAX_DIM] = {0, 0, 0};
    separation_vector(c->pos + c->st.active * c->D, c->pos + target * c->D, c->D, c->L, sep);
    double c1 = c->prog.pair_use_charge ? c->charge[c->st.active] : 1.0;
    double c2 = c->prog.pair_use_charge ? c->charge[target] : 1.0;
    double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter,
                          ECMC_SLOT(ECMC_SLOT_PAIR_TIME, target / c->npr), (uint32_t)(target % c->npr));
    double dt = pot_displacement(&c->pair_bound, c->st.direction, c->prog.speed, sep, c->D, c1, c2,
                                 rng_expovariate(u, c->prog.beta));
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, dt);
    return cand;
}

/* A two-leaf factor between the active leaf and a leaf of another object (non-local factor type map entry), handled by
 * a TwoLeafUnitEventHandler (two_leaf_unit_event_handler.py:105-138) */
static candidate factor_pair_candidate(OrcChain *c, int target) {
    candidate cand;
    cand.kind = ECMC_EVENT_FACTOR_PAIR; cand.target = target; cand.target_cell = -1; cand.rate = 0.0;
    double sep[ECMC_MAX_DIM] = {0, 0, 0};
    separation_vector(c->pos + c->st.active * c->D, c->pos + target * c->D, c->D, c->L, sep);
    double dU = 0.0;
    if (pot_needs_potential_change(c->inter_pot.kind)) {
        double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_FACTOR_TIME, target), 0);
        dU = rng_expovariate(u, c->prog.beta);
    }
    double dt = pot_displacement(&c->inter_pot, c->st.direction, c->prog.speed, sep, c->D, 1.0, 1.0, dU);
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, dt);
    return cand;
}

/* the derivative triple of the bending factor for the three units of the active object, the active leaf optionally
 * displaced (event_handler_with_bounding_potential.py:282-332; _get_separations of
 * fixed_separations_event_handler_with_piecewise_constant_bounding_potential.py:189-210) */
static void bending_triple(OrcChain *c, double advance, double out[3]) {
    int root = c->st.active / c->npr;
    double positions[3][ECMC_MAX_DIM];
    for (int i = 0; i < 3; i++) {
        int leaf = root * c->npr + c->prog.bending_children[i];
        for (int d = 0; d < c->D; d++) positions[i][d] = c->pos[leaf * c->D + d];
        if (leaf == c->st.active && advance != 0.0)
            for (int d = 0; d < c->D; d++) {
                double v = d == c->st.direction ? c->prog.speed : 0.0;
                positions[i][d] = correct_position_entry(positions[i][d] + v * advance, c->L);
            }
    }
    const int32_t *s = c->prog.bending_separations;
    double s1[ECMC_MAX_DIM] = {0, 0, 0}, s2[ECMC_MAX_DIM] = {0, 0, 0};
    separation_vector(positions[s[0]], positions[s[1]], c->D, c->L, s1);
    separation_vector(positions[s[2]], positions[s[3]], c->D, c->L, s2);
    bending_derivative(c->bending_pot.p0, c->bending_pot.p1, c->st.direction, c->prog.speed, s1, s2, c->D, out);
}
static int bending_active_index(const OrcChain *c) {
    int child = c->st.active % c->npr;
    for (int i = 0; i < 3; i++) if (c->prog.bending_children[i] == child) return i;
    return -1;
}
/* FixedSeparationsEventHandlerWithPiecewiseConstantBoundingPotential.send_event_time (:113-141): cand.rate carries the
 * bounding event rate, or a negative number for None */
static candidate bending_candidate(OrcChain *c) {
    candidate cand;
    cand.kind = ECMC_EVENT_BENDING; cand.target = -1; cand.target_cell = -1;
    int index = bending_active_index(c);
    double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_BENDING_TIME, 0), 0);
    double potential_change = rng_expovariate(u, c->prog.beta);
    double one[3], two[3];
    bending_triple(c, 0.0, one);
    bending_triple(c, c->prog.bending_max_displacement, two);
    double constant = (one[index] > two[index] ? one[index] : two[index]) + c->prog.bending_offset; /* max(a, b) */
    double dt;
    if (constant <= 0.0) { cand.rate = -1.0; dt = c->prog.bending_max_displacement; }
    else if (potential_change / constant < c->prog.bending_max_displacement) { cand.rate = constant; dt = potential_change / constant; }
    else { cand.rate = -1.0; dt = c->prog.bending_max_displacement; }
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, dt);
    return cand;
}

/* CellBoundaryEventHandler.send_event_time for the root unit (cell_level = 1): it moves with velocity * weight */
static candidate root_boundary_candidate(OrcChain *c, double *boundary_out) {
    candidate cand;
    cand.kind = ECMC_EVENT_CELL_BOUNDARY; cand.target = -1; cand.rate = 0.0;
    int dir = c->st.direction;
    const double *pr = c->root_pos + (c->st.active / c->npr) * c->D;
    int cell = position_to_cell(&c->cells, pr);
    int neighbor = neighbor_cell_positive(&c->cells, cell, dir);
    double neighbor_boundary = c->cells.cell_min[neighbor * c->D + dir];
    double separation = neighbor_boundary - pr[dir];
    if (separation < 0.0) separation = separation + c->L;
    double velocity = c->prog.speed * (1.0 / c->npr);
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, separation / velocity);
    cand.target_cell = neighbor;
    *boundary_out = neighbor_boundary;
    return cand;
}

/* EventHandlerWithBoundingPotential._fill_lifting (event_handler_with_bounding_potential.py:170-220) followed by the
 * lifting scheme: returns the new active leaf. target_derivatives holds -derivative(active, target_k) on entry. */
static int composite_lifting(OrcChain *c, int target_root, double active_derivative, double *target_derivatives,
                             uint32_t *draw) {
    int local_root = c->st.active / c->npr;
    double local_derivatives[4] = {0, 0, 0, 0};
    for (int i = 0; i < c->npr; i++) {
        int local = local_root * c->npr + i;
        if (local == c->st.active) { local_derivatives[i] = active_derivative; continue; }
        for (int j = 0; j < c->npr; j++) {
            int target = target_root * c->npr + j;
            double sep[ECMC_MAX_DIM] = {0, 0, 0};
            separation_vector(c->pos + local * c->D, c->pos + target * c->D, c->D, c->L, sep);
            double c1 = c->prog.pair_use_charge ? c->charge[local] : 1.0;
            double c2 = c->prog.pair_use_charge ? c->charge[target] : 1.0;
            double pairwise = pot_derivative(&c->pair_pot, c->st.direction, c->prog.speed, sep, c->D, c1, c2);
            local_derivatives[i] += pairwise;
            target_derivatives[j] -= pairwise;
        }
    }
    olifting lift;
    lifting_reset(&lift);
    for (int pass = 0; pass < 2; pass++) {
        int local_now = (local_root < target_root) == (pass == 0);
        for (int i = 0; i < c->npr; i++) {
            if (local_now)
                lifting_insert(&lift, local_derivatives[i], local_root * c->npr + i, local_root * c->npr + i == c->st.active, c, draw);
            else
                lifting_insert(&lift, target_derivatives[i], target_root * c->npr + i, 0, c, draw);
        }
    }
    return lifting_get(&lift, c->prog.composite_lifting, c, draw);
}

/* TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential.send_event_time
 * (two_leaf_unit_event_handler_with_piecewise_constant_bounding_potential.py:103-131 on
 * _displacement_from_piecewise_constant_bounding_potential, event_handler_with_bounding_potential.py:282-332): the active
 * leaf against leaf `target` of another object; rate < 0 stands for "bounding event rate None" */
static candidate piecewise_pair_candidate(OrcChain *c, int target) {
    candidate cand;
    cand.kind = ECMC_EVENT_FACTOR_PAIR; cand.target = target; cand.target_cell = -1;
    const int D = c->D, dir = c->st.direction;
    const double *pa = c->pos + c->st.active * D, *pt = c->pos + target * D;
    double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_FACTOR_TIME, target), 0);
    double potential_change = rng_expovariate(u, c->prog.beta);
    double sep[ECMC_MAX_DIM] = {0, 0, 0}, moved[ECMC_MAX_DIM] = {0, 0, 0};
    separation_vector(pa, pt, D, c->L, sep);
    double one = pot_derivative(&c->inter_pot, dir, c->prog.speed, sep, D, 1.0, 1.0);
    for (int d = 0; d < D; d++) {
        double v = d == dir ? c->prog.speed : 0.0;
        moved[d] = correct_position_entry(pa[d] + v * c->prog.inter_bound_max_displacement, c->L);
    }
    separation_vector(moved, pt, D, c->L, sep);
    double two = pot_derivative(&c->inter_pot, dir, c->prog.speed, sep, D, 1.0, 1.0);
    double constant = (one > two ? one : two) + c->prog.inter_bound_offset; /* max(a, b) */
    double dt;
    if (constant <= 0.0) { cand.rate = -1.0; dt = c->prog.inter_bound_max_displacement; }
    else if (potential_change / constant < c->prog.inter_bound_max_displacement) { cand.rate = constant; dt = potential_change / constant; }
    else { cand.rate = -1.0; dt = c->prog.inter_bound_max_displacement; }
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, dt);
    return cand;
}

/* TwoLeafUnitCellBoundingPotentialEventHandler.send_event_time (two_leaf_unit_cell_bounding_potential_event_handler.py:
 * 137-177) for the active leaf and leaf `target` in a cell that is not nearby; chargeless here. The draw is double 1 of
 * the factor-time slot of the target leaf (the pair-time slots belong to the composite-object handlers). */
static candidate leaf_cell_bounding_candidate(OrcChain *c, int target) {
    candidate cand;
    cand.kind = ECMC_EVENT_CELL_BOUNDING; cand.target = target;
    int dir = c->st.direction;
    int active_cell = position_to_cell(&c->cells, c->pos + c->st.active * c->D);
    int target_cell = position_to_cell(&c->cells, c->pos + target * c->D);
    int relative_cell = cells_relative(&c->cells, target_cell, active_cell);
    cand.target_cell = relative_cell;
    cand.rate = c->bounds[(relative_cell * c->D + dir) * 2 + 0];
    double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_FACTOR_TIME, target), 1);
    double dU = rng_expovariate(u, c->prog.beta);
    double displacement = cand.rate > 0 ? dU / cand.rate : ORC_INF;
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, displacement / c->prog.speed);
    return cand;
}

/* SingleActiveCellOccupancy.update (:149-203) for a cell system that stores one kind of leaf only: the old active leaf
 * goes back into its cell if it is of that kind, the new one leaves its cell if it is */
static void leaf_cells_occupancy_update(OrcChain *c, int new_active) {
    const int kind = c->prog.cell_child - 1, npr = c->npr;
    const int old_active = c->st.active;
    const int old_relevant = old_active % npr == kind, new_relevant = new_active % npr == kind;
    if (new_active != old_active) {
        if (old_relevant) occ_insert(c, c->st.active_cell, old_active);
        c->st.active_cell = 0;
        if (new_relevant) {
            c->st.active_cell = position_to_cell(&c->cells, c->pos + new_active * c->D);
            occ_remove(c, c->st.active_cell, new_active);
        }
    } else if (new_relevant) {
        c->st.active_cell = position_to_cell(&c->cells, c->pos + new_active * c->D);
    }
    c->st.active = new_active;
}

/* SingleActiveCellOccupancy.update for cell_level = 1: the active unit on the cell level is the root */
static void molecule_occupancy_update(OrcChain *c, int new_active) {
    if (c->prog.cell_child) { leaf_cells_occupancy_update(c, new_active); return; }
    int old_root = c->st.active / c->npr, new_root = new_active / c->npr;
    if (new_root != old_root) {
        occ_insert(c, c->st.active_cell, old_root);
        c->st.active_cell = position_to_cell(&c->cells, c->root_pos + new_root * c->D);
        occ_remove(c, c->st.active_cell, new_root);
    } else {
        c->st.active_cell = position_to_cell(&c->cells, c->root_pos + new_root * c->D);
    }
    c->st.active = new_active;
}

static int is_leaf_kind(int kind) {
    return kind == ECMC_EVENT_PAIR || kind == ECMC_EVENT_CELL_BOUNDING || kind == ECMC_EVENT_BOND ||
           kind == ECMC_EVENT_FACTOR_PAIR;
}

static int root_mode_step(OrcChain *c, otime until, EcmcEventRecord *rec);

static int molecule_step(OrcChain *c, otime until, EcmcEventRecord *rec) {
    if (c->prog.root_mode && c->st.mode == 1) return root_mode_step(c, until, rec);
    candidate best;
    int n_cand = 0;
    double boundary_position = 0.0;
    best.kind = ECMC_EVENT_NONE; best.t.q = ORC_INF; best.t.r = ORC_INF; best.target = -1; best.target_cell = -1;
    best.rate = 0.0;
    int was_pending = c->st.pending_kind != ECMC_EVENT_NONE;
    int npr = c->npr, D = c->D, dir = c->st.direction;
    int active_root = c->st.active / npr, active_child = c->st.active % npr;
    candidate factor_best; /* the earliest leaf-level factor candidate of this iteration (computed or kept) */
    int factor_from_kept = 0, best_from_kept = 0;
    factor_best.kind = ECMC_EVENT_NONE; factor_best.t.q = ORC_INF; factor_best.t.r = ORC_INF; factor_best.target = -1;
    factor_best.target_cell = -1; factor_best.rate = 0.0;
    if (was_pending) {
        best.kind = c->st.pending_kind;
        best.t.q = c->st.pending_q; best.t.r = c->st.pending_r;
        best.rate = c->st.pending_rate;
        if (is_leaf_kind(best.kind)) best.target = c->st.pending_target; else best.target_cell = c->st.pending_target;
        if (best.kind == ECMC_EVENT_CELL_BOUNDARY) boundary_position = c->cells.cell_min[best.target_cell * D + dir];
    } else {
#define CONSIDER(cand) do { if (!isinf((cand).t.q)) { n_cand++; if (lt_candidate(&(cand), &best)) { best = (cand); best_from_kept = 0; } } } while (0)
        int nearby[343];
        int nn = nearby_cells(&c->cells, c->st.active_cell, nearby);
        const int leaf_cells = c->prog.cell_child != 0;
        const int cell_leaf_active = leaf_cells && active_child == c->prog.cell_child - 1;
        if (leaf_cells) {
            /* the composite-object pairs come from the factor type map and belong to the handlers that a cell-boundary
             * event leaves running: below, with the bonds and the bending factor */
        } else if (c->prog.pair_handler == ECMC_PAIR_TWO_COMPOSITE_SUMMED_BOUNDING) {
            for (int i = 0; i < nn; i++) {
                int t = c->occ[nearby[i]];
                if (t < 0) continue;
                candidate cand = composite_pair_candidate(c, t);
                CONSIDER(cand);
            }
            for (int s = 0; s < c->n_surplus; s++) {
                candidate cand = composite_pair_candidate(c, c->surplus[s]);
                CONSIDER(cand);
            }
        } else if (c->prog.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING) {
            /* one bounded handler per leaf of every other object */
            for (int i = 0; i < nn + c->n_surplus; i++) {
                int t = i < nn ? c->occ[nearby[i]] : c->surplus[i - nn];
                if (t < 0) continue;
                for (int k = 0; k < npr; k++) {
                    candidate cand = leaf_pair_candidate(c, t * npr + k);
                    CONSIDER(cand);
                }
            }
        }
        /* the leaf-level factors: recomputed, unless a cell-boundary event left their handlers running */
        if (c->st.kept_kind != ECMC_EVENT_NONE) {
            if (c->st.kept_kind > 0) {
                factor_best.kind = c->st.kept_kind;
                factor_best.t.q = c->st.kept_q; factor_best.t.r = c->st.kept_r;
                factor_best.rate = c->st.kept_rate;
                factor_best.target = c->st.kept_target;
                factor_from_kept = 1;
            }
        } else {
#define CONSIDER_FACTOR(cand) do { if (!isinf((cand).t.q)) { n_cand++; if (lt_candidate(&(cand), &factor_best)) factor_best = (cand); } } while (0)
            if (leaf_cells)
                for (int r = 0; r < c->N / npr; r++) {
                    if (r == active_root) continue;
                    candidate cand = composite_pair_candidate(c, r);
                    CONSIDER_FACTOR(cand);
                }
            for (int b = 0; b < c->prog.n_bonds; b++) {
                int partner = -1;
                if (c->prog.bonds[b][0] == active_child) partner = c->prog.bonds[b][1];
                else if (c->prog.bonds[b][1] == active_child) partner = c->prog.bonds[b][0];
                if (partner < 0) continue;
                candidate cand = bond_candidate(c, active_root * npr + partner);
                CONSIDER_FACTOR(cand);
            }
            for (int f = 0; f < c->prog.n_inter_factors && !leaf_cells; f++) {
                if (c->prog.inter_factors[f][0] != active_child) continue;
                for (int r = 0; r < c->N / npr; r++) {
                    if (r == active_root) continue;
                    candidate cand = factor_pair_candidate(c, r * npr + c->prog.inter_factors[f][1]);
                    CONSIDER_FACTOR(cand);
                }
            }
            if (c->prog.bending_enabled && bending_active_index(c) >= 0) {
                candidate cand = bending_candidate(c);
                CONSIDER_FACTOR(cand);
            }
#undef CONSIDER_FACTOR
        }
        if (lt_candidate(&factor_best, &best)) { best = factor_best; best_from_kept = factor_from_kept; }
        if (leaf_cells) {
            /* found through the cells of the active leaf, if it is of the stored kind (the occupancy has no active cell
             * otherwise): ExcludedCellsTagger + SurplusCellsTagger -> the piecewise-constant-bound handler,
             * CellBoundingPotentialTagger -> the cell-bounding handler, CellBoundaryTagger */
            if (cell_leaf_active) {
                for (int i = 0; i < nn; i++) {
                    int t = c->occ[nearby[i]];
                    if (t < 0) continue;
                    candidate cand = piecewise_pair_candidate(c, t);
                    CONSIDER(cand);
                }
                for (int sidx = 0; sidx < c->n_surplus; sidx++) {
                    candidate cand = piecewise_pair_candidate(c, c->surplus[sidx]);
                    CONSIDER(cand);
                }
                for (int cell = 0; cell < c->cells.n_cells && c->prog.veto_enabled == ECMC_FAR_CELL_BOUNDING; cell++) {
                    int t = c->occ[cell];
                    if (t < 0) continue;
                    int is_near = 0;
                    for (int i = 0; i < nn; i++) if (nearby[i] == cell) { is_near = 1; break; }
                    if (is_near) continue;
                    candidate cand = leaf_cell_bounding_candidate(c, t);
                    CONSIDER(cand);
                }
                candidate cand = boundary_candidate(c, &boundary_position);
                n_cand++;
                if (lt_candidate(&cand, &best)) { best = cand; best_from_kept = 0; }
            }
        } else if (c->prog.veto_enabled == ECMC_FAR_CELL_VETO) {
            candidate cand = veto_candidate(c);
            CONSIDER(cand);
        } else if (c->prog.veto_enabled == ECMC_FAR_CELL_BOUNDING) {
            /* CellBoundingPotentialTagger (cell_bounding_potential_tagger.py:150-155): every occupied cell that is not
             * nearby the active cell */
            for (int cell = 0; cell < c->cells.n_cells; cell++) {
                int t = c->occ[cell];
                if (t < 0) continue;
                int is_near = 0;
                for (int i = 0; i < nn; i++) if (nearby[i] == cell) { is_near = 1; break; }
                if (is_near) continue;
                candidate cand = composite_cell_bounding_candidate(c, t);
                CONSIDER(cand);
            }
        }
        if (!c->prog.no_cells && !leaf_cells) {
            candidate cand = root_boundary_candidate(c, &boundary_position);
            n_cand++;
            if (lt_candidate(&cand, &best)) { best = cand; best_from_kept = 0; }
        }
#undef CONSIDER
    }
    {
        candidate interaction = best;
        candidate cand;
        cand.kind = ECMC_EVENT_END_OF_CHAIN; cand.target = c->st.eoc_next_active; cand.target_cell = -1; cand.rate = 0.0;
        cand.t.q = c->st.eoc_q; cand.t.r = c->st.eoc_r;
        n_cand++;
        if (lt_candidate(&cand, &best)) best = cand;
        if (c->prog.root_mode) {
            /* the leaf-to-root RootLeafUnitActiveSwitcher, in the scheduler since the last switch */
            cand.kind = ECMC_EVENT_SWITCH; cand.target = -1;
            cand.t.q = c->st.switch_q; cand.t.r = c->st.switch_r;
            n_cand++;
            if (lt_candidate(&cand, &best)) best = cand;
        }
        if (!time_lt(best.t, until)) {
            c->st.pending_kind = interaction.kind;
            c->st.pending_q = interaction.t.q; c->st.pending_r = interaction.t.r;
            c->st.pending_rate = interaction.rate;
            c->st.pending_target = is_leaf_kind(interaction.kind) ? interaction.target : interaction.target_cell;
            if (!was_pending) {
                c->st.pending_position = c->pos[c->st.active * D + dir];
                c->st.pending_root_position = c->root_pos[active_root * D + dir];
                c->st.pending_stamp_q = c->st.time_q;
                c->st.pending_stamp_r = c->st.time_r;
                if (best_from_kept) { /* the kept factor handler's in-state is older still */
                    c->st.pending_position = c->st.kept_position;
                    c->st.pending_root_position = c->st.kept_root_position;
                    c->st.pending_stamp_q = c->st.kept_stamp_q;
                    c->st.pending_stamp_r = c->st.kept_stamp_r;
                }
            }
            return 0;
        }
    }
    c->st.pending_kind = ECMC_EVENT_NONE;
    if (was_pending && best.kind != ECMC_EVENT_END_OF_CHAIN && best.kind != ECMC_EVENT_SWITCH) {
        c->pos[c->st.active * D + dir] = c->st.pending_position;
        c->root_pos[active_root * D + dir] = c->st.pending_root_position;
        c->st.time_q = c->st.pending_stamp_q;
        c->st.time_r = c->st.pending_stamp_r;
    } else if (best_from_kept && best.kind != ECMC_EVENT_END_OF_CHAIN) {
        /* the kept handler computes its out-state from the in-state it received before the cell-boundary event(s) */
        c->pos[c->st.active * D + dir] = c->st.kept_position;
        c->root_pos[active_root * D + dir] = c->st.kept_root_position;
        c->st.time_q = c->st.kept_stamp_q;
        c->st.time_r = c->st.kept_stamp_r;
    }
    /* which handlers survive this event: a cell-boundary event leaves the leaf-level factors running */
    if (best.kind == ECMC_EVENT_CELL_BOUNDARY && c->prog.boundary_keeps_factors) {
        if (c->st.kept_kind == ECMC_EVENT_NONE && !was_pending) {
            c->st.kept_kind = factor_best.kind == ECMC_EVENT_NONE ? -1 : factor_best.kind;
            c->st.kept_target = factor_best.target;
            c->st.kept_q = factor_best.t.q; c->st.kept_r = factor_best.t.r;
            c->st.kept_rate = factor_best.rate;
            c->st.kept_position = c->pos[c->st.active * D + dir];
            c->st.kept_root_position = c->root_pos[active_root * D + dir];
            c->st.kept_stamp_q = c->st.time_q; c->st.kept_stamp_r = c->st.time_r;
        }
    } else {
        c->st.kept_kind = ECMC_EVENT_NONE;
    }

    int old_active = c->st.active, new_active = old_active, accepted = 0, rec_target = -1;
    uint32_t draw = 0;
    time_slice_active(c, best.t);
    const double *pa = c->pos + old_active * D;
    switch (best.kind) {
    case ECMC_EVENT_PAIR:
    case ECMC_EVENT_CELL_BOUNDING:
    case ECMC_EVENT_CELL_VETO: {
        /* two_composite_object_summed_bounding_potential_event_handler.py:158-202 /
         * composite_object_cell_veto_event_handler.py:110-162 (+ mediator.py:265-292 for the occupant of the cell) /
         * two_composite_object_cell_bounding_potential_event_handler.py:198-246: like the cell veto with a known target
         * object; the stored rate is a rate per length, the bounding potential's derivative is that times the speed */
        int far = best.kind == ECMC_EVENT_CELL_BOUNDING;
        int veto = best.kind == ECMC_EVENT_CELL_VETO || far;
        if (far && c->prog.cell_child) {
            /* TwoLeafUnitCellBoundingPotentialEventHandler.send_out_state (:179-211) between two leaves: the stored rate
             * times the speed is the bounding event rate, confirmed against the real potential
             * (event_handler_with_bounding_potential.py:75-101); recorded with the target's object */
            rec_target = best.target / npr;
            c->stats.pair_events++;
            double sep[ECMC_MAX_DIM] = {0, 0, 0};
            separation_vector(pa, c->pos + best.target * D, D, c->L, sep);
            double bounding_rate = best.rate * c->prog.speed;
            double real = pot_derivative(&c->veto_pot, dir, c->prog.speed, sep, D, 1.0, 1.0);
            if (real > 0) {
                if (bounding_rate < real) c->stats.bound_violations++;
                double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), draw++);
                if (0 + (bounding_rate - 0) * u < real) { accepted = 1; new_active = best.target; }
            }
            break;
        }
        if (!veto && c->prog.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT_BOUNDING) {
            /* TwoLeafUnitBoundingPotentialEventHandler.send_out_state (:148-168) +
             * _calculate_out_state_of_two_leaf_unit_bounding_potential (event_handler_with_bounding_potential.py:75-101) */
            rec_target = best.target;
            c->stats.pair_events++;
            double sep[ECMC_MAX_DIM] = {0, 0, 0};
            separation_vector(pa, c->pos + best.target * D, D, c->L, sep);
            double c1 = c->prog.pair_use_charge ? c->charge[old_active] : 1.0;
            double c2 = c->prog.pair_use_charge ? c->charge[best.target] : 1.0;
            double bounding_rate = pot_derivative(&c->pair_bound, dir, c->prog.speed, sep, D, c1, c2);
            double real = pot_derivative(&c->pair_pot, dir, c->prog.speed, sep, D, c1, c2);
            if (real > 0) {
                if (bounding_rate < real) c->stats.bound_violations++;
                double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), draw++);
                if (0 + (bounding_rate - 0) * u < real) { accepted = 1; new_active = best.target; }
            }
            break;
        }
        int target_root = (veto && !far) ? c->occ[best.target_cell] : best.target;
        rec_target = target_root;
        if (veto && !far) c->stats.veto_events++; else c->stats.pair_events++;
        if (target_root < 0) break;
        double bounding_rate = far ? best.rate * c->prog.speed : (veto ? best.rate : 0.0);
        double factor_derivative = 0.0;
        double target_derivatives[4] = {0, 0, 0, 0};
        const opotential *real = veto ? &c->veto_pot : &c->pair_pot;
        int use_charge = veto ? c->prog.veto_use_charge : c->prog.pair_use_charge;
        for (int k = 0; k < npr; k++) {
            int t = target_root * npr + k;
            double sep[ECMC_MAX_DIM] = {0, 0, 0};
            separation_vector(pa, c->pos + t * D, D, c->L, sep);
            double c1 = use_charge ? c->charge[old_active] : 1.0, c2 = use_charge ? c->charge[t] : 1.0;
            if (!veto) {
                double b = pot_derivative(&c->pair_bound, dir, c->prog.speed, sep, D, c1, c2);
                bounding_rate += b > 0.0 ? b : 0.0; /* max(0.0, b) */
            }
            double pairwise = pot_derivative(real, dir, c->prog.speed, sep, D, c1, c2);
            factor_derivative += pairwise;
            target_derivatives[k] -= pairwise;
        }
        double event_rate = factor_derivative > 0.0 ? factor_derivative : 0.0; /* max(0.0, factor_derivative) */
        if (bounding_rate < event_rate) c->stats.bound_violations++;
        double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), draw++);
        if (event_rate <= 0.0 + (bounding_rate - 0.0) * u) break;
        /* the summed-bounding handler passes the clipped event rate, the cell-veto handler the factor derivative */
        new_active = composite_lifting(c, target_root, veto ? factor_derivative : event_rate, target_derivatives, &draw);
        accepted = 1;
        if (veto && !far) c->stats.veto_accepted++;
        break;
    }
    case ECMC_EVENT_BOND:
    case ECMC_EVENT_FACTOR_PAIR:
        rec_target = best.target;
        if (best.kind == ECMC_EVENT_FACTOR_PAIR && c->prog.cell_child) {
            /* TwoLeafUnitEventHandlerWithPiecewiseConstantBoundingPotential.send_out_state (:133-152) */
            c->stats.factor_pair_events++;
            if (best.rate < 0.0) break; /* bounding event rate None */
            double sep[ECMC_MAX_DIM] = {0, 0, 0};
            separation_vector(pa, c->pos + best.target * D, D, c->L, sep);
            double real = pot_derivative(&c->inter_pot, dir, c->prog.speed, sep, D, 1.0, 1.0);
            if (real > 0) {
                if (best.rate < real) c->stats.bound_violations++;
                double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), draw++);
                if (0.0 + (best.rate - 0.0) * u < real) { accepted = 1; new_active = best.target; }
            }
            break;
        }
        accepted = 1;
        new_active = best.target;
        if (best.kind == ECMC_EVENT_BOND) c->stats.bond_events++; else c->stats.factor_pair_events++;
        break;
    case ECMC_EVENT_BENDING: {
        /* fixed_separations_event_handler_with_piecewise_constant_bounding_potential.py:143-187 */
        c->stats.bond_events++;
        if (best.rate < 0.0) break; /* bounding event rate None */
        double derivatives[3];
        bending_triple(c, 0.0, derivatives);
        int index = bending_active_index(c);
        if (derivatives[index] > 0) {
            if (best.rate < derivatives[index]) c->stats.bound_violations++;
            double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), draw++);
            if (0.0 + (best.rate - 0.0) * u < derivatives[index]) {
                olifting lift;
                lifting_reset(&lift);
                for (int i = 0; i < 3; i++)
                    lifting_insert(&lift, derivatives[i], active_root * npr + c->prog.bending_children[i], i == index, c, &draw);
                new_active = lifting_get(&lift, c->prog.bending_lifting, c, &draw);
                accepted = 1;
            }
        }
        break;
    }
    case ECMC_EVENT_CELL_BOUNDARY:
        /* the unit on the cell level lands on the boundary (cell_boundary_event_handler.py:158-173): the root unit, or
         * the active leaf of a cell system of leaves */
        if (c->prog.cell_child) c->pos[old_active * D + dir] = boundary_position;
        else c->root_pos[active_root * D + dir] = boundary_position;
        c->stats.boundary_events++;
        break;
    case ECMC_EVENT_END_OF_CHAIN:
        new_active = c->st.eoc_next_active;
        rec_target = new_active;
        accepted = 1;
        c->stats.end_of_chain_events++;
        if (c->prog.root_mode) { c->st.eoc_last_q = best.t.q; c->st.eoc_last_r = best.t.r; }
        break;
    case ECMC_EVENT_SWITCH:
        /* RootLeafUnitActiveSwitcher._send_out_state_root_unit_active (root_leaf_unit_active_switcher.py:171-208): the
         * other leaves of the object take the velocity and the time stamp of the active leaf, the root unit ends with the
         * full velocity; recorded with the first leaf of the object as the active one */
        new_active = active_root * npr;
        accepted = 1;
        c->st.mode = 1;
        break;
    default: break;
    }
    if (new_active < 0) { c->stats.capacity_errors++; new_active = old_active; }
    if (rec) {
        memset(rec, 0, sizeof(*rec));
        rec->kind = best.kind;
        rec->target = rec_target;
        rec->target_cell = best.target_cell;
        rec->accepted = (best.kind == ECMC_EVENT_END_OF_CHAIN || best.kind == ECMC_EVENT_SWITCH) ? 1 : (new_active != old_active);
        rec->n_candidates = n_cand;
        rec->time_q = best.t.q; rec->time_r = best.t.r;
        for (int d = 0; d < D; d++) rec->active_pos[d] = c->pos[old_active * D + d];
    }
    (void)accepted;
    c->st.event_counter++;
    c->stats.events++;
    c->stats.candidates += (uint64_t)n_cand;
    if (best.kind == ECMC_EVENT_END_OF_CHAIN) c->st.direction = (c->st.direction + 1) % D;
    molecule_occupancy_update(c, new_active);
    if (best.kind == ECMC_EVENT_SWITCH) {
        /* the root-to-leaf switcher is created with the root unit's time stamp, the end of chain is re-created */
        otime t = time_add(best.t, c->prog.switch_chain_length[1]);
        c->st.switch_q = t.q; c->st.switch_r = t.r;
    }
    if (best.kind == ECMC_EVENT_END_OF_CHAIN || best.kind == ECMC_EVENT_SWITCH) schedule_end_of_chain(c);
    if (rec) { rec->new_active = c->st.active; rec->new_direction = c->st.direction; rec->mode = c->st.mode; }
    return 1;
}

/* ---- the root unit of an object is the independent active unit (EcmcProgram.root_mode, dipoles/dipole_motion.ini):
 * root and leaves move with the full velocity and carry one time stamp. `active` is the first leaf of the object. */
static void time_slice_object(OrcChain *c, otime event_time) {
    /* BasicEventHandler._time_slice_unit for the root unit and every leaf (abstracts.py:89-107) */
    const int D = c->D, root = c->st.active / c->npr;
    otime stamp = {c->st.time_q, c->st.time_r};
    double dt = time_sub(event_time, stamp);
    for (int d = 0; d < D; d++) {
        double v = d == c->st.direction ? c->prog.speed : 0.0;
        for (int k = 0; k < c->npr; k++) {
            double *p = c->pos + (root * c->npr + k) * D;
            p[d] = correct_position_entry(p[d] + v * dt, c->L);
        }
        double *pr = c->root_pos + root * D;
        pr[d] = correct_position_entry(pr[d] + v * dt, c->L);
    }
    c->st.time_q = event_time.q; c->st.time_r = event_time.r;
}

static int root_mode_step(OrcChain *c, otime until, EcmcEventRecord *rec) {
    const int npr = c->npr, D = c->D, dir = c->st.direction, n_roots = c->N / npr;
    const int active_root = c->st.active / npr;
    candidate best;
    int n_cand = 0;
    best.kind = ECMC_EVENT_NONE; best.t.q = ORC_INF; best.t.r = ORC_INF; best.target = -1; best.target_cell = -1;
    best.rate = 0.0;
    const int was_pending = c->st.pending_kind != ECMC_EVENT_NONE;
    otime now = {c->st.time_q, c->st.time_r};
    if (was_pending) {
        best.kind = c->st.pending_kind;
        best.t.q = c->st.pending_q; best.t.r = c->st.pending_r;
        best.target = c->st.pending_target;
    } else {
        for (int t = 0; t < n_roots; t++) {
            if (t == active_root) continue;
            /* RootUnitActiveTwoCompositeObjectSummedBoundingPotentialEventHandler.send_event_time (:116-152): minimum over
             * the local leaf units x the target leaf units (both sorted by identifier) of the bounding potential's
             * displacement, one expovariate each in that order */
            candidate cand;
            cand.kind = ECMC_EVENT_PAIR; cand.target = t; cand.target_cell = -1; cand.rate = 0.0;
            double earliest = ORC_INF;
            for (int i = 0; i < npr; i++)
                for (int j = 0; j < npr; j++) {
                    int local = active_root * npr + i, target = t * npr + j;
                    double sep[ECMC_MAX_DIM] = {0, 0, 0};
                    separation_vector(c->pos + local * D, c->pos + target * D, D, c->L, sep);
                    double c1 = c->prog.pair_use_charge ? c->charge[local] : 1.0;
                    double c2 = c->prog.pair_use_charge ? c->charge[target] : 1.0;
                    double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter,
                                          ECMC_SLOT(ECMC_SLOT_PAIR_TIME, t), (uint32_t)(i * npr + j));
                    double dt = pot_displacement(&c->pair_bound, dir, c->prog.speed, sep, D, c1, c2,
                                                 rng_expovariate(u, c->prog.beta));
                    if (dt < earliest) earliest = dt;
                }
            cand.t = time_add(now, earliest);
            if (!isinf(cand.t.q)) { n_cand++; if (lt_candidate(&cand, &best)) best = cand; }
            /* RootUnitActiveTwoLeafUnitEventHandler (send_event_time of TwoLeafUnitEventHandler,
             * two_leaf_unit_event_handler.py:105-138, for the moving leaf of the factor): one per factor type map entry */
            for (int f = 0; f < c->prog.n_inter_factors; f++) {
                int local = active_root * npr + c->prog.inter_factors[f][0];
                int target = t * npr + c->prog.inter_factors[f][1];
                double sep[ECMC_MAX_DIM] = {0, 0, 0};
                separation_vector(c->pos + local * D, c->pos + target * D, D, c->L, sep);
                double dU = 0.0;
                if (pot_needs_potential_change(c->inter_pot.kind)) {
                    double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter,
                                          ECMC_SLOT(ECMC_SLOT_FACTOR_TIME, target), (uint32_t)c->prog.inter_factors[f][0]);
                    dU = rng_expovariate(u, c->prog.beta);
                }
                candidate fc;
                fc.kind = ECMC_EVENT_FACTOR_PAIR; fc.target = target; fc.target_cell = -1; fc.rate = 0.0;
                fc.t = time_add(now, pot_displacement(&c->inter_pot, dir, c->prog.speed, sep, D, 1.0, 1.0, dU));
                if (!isinf(fc.t.q)) { n_cand++; if (lt_candidate(&fc, &best)) best = fc; }
            }
        }
    }
    {
        candidate interaction = best;
        candidate cand;
        cand.kind = ECMC_EVENT_END_OF_CHAIN; cand.target = c->st.eoc_next_active; cand.target_cell = -1; cand.rate = 0.0;
        cand.t.q = c->st.eoc_q; cand.t.r = c->st.eoc_r;
        n_cand++;
        if (lt_candidate(&cand, &best)) best = cand;
        cand.kind = ECMC_EVENT_SWITCH; cand.target = -1;
        cand.t.q = c->st.switch_q; cand.t.r = c->st.switch_r;
        n_cand++;
        if (lt_candidate(&cand, &best)) best = cand;
        if (!time_lt(best.t, until)) {
            c->st.pending_kind = interaction.kind;
            c->st.pending_q = interaction.t.q; c->st.pending_r = interaction.t.r;
            c->st.pending_rate = 0.0;
            c->st.pending_target = interaction.target;
            if (!was_pending) {
                /* the in-state of the handler that stays in the scheduler: the coordinates of the two leaves */
                c->st.pending_position = c->pos[(active_root * npr) * D + dir];
                c->st.pending_position_y = c->pos[(active_root * npr + 1) * D + dir];
                c->st.pending_root_position = c->root_pos[active_root * D + dir];
                c->st.pending_stamp_q = c->st.time_q;
                c->st.pending_stamp_r = c->st.time_r;
            }
            return 0;
        }
    }
    c->st.pending_kind = ECMC_EVENT_NONE;
    /* The root-unit-active handlers get a fresh copy of the objects for their out-state (mediator.py:
     * get_arguments_composite_objects_lifting) and time-slice THAT to the event time; only the confirmation of the
     * summed-bounding handler uses the leaf units it time-sliced in send_event_time (:149, :154-168). */
    double in_state[2][ECMC_MAX_DIM];
    for (int k = 0; k < npr; k++)
        for (int d = 0; d < D; d++) in_state[k][d] = c->pos[(active_root * npr + k) * D + d];
    otime in_stamp = now;
    if (was_pending) {
        in_state[0][dir] = c->st.pending_position;
        in_state[1][dir] = c->st.pending_position_y;
        in_stamp.q = c->st.pending_stamp_q; in_stamp.r = c->st.pending_stamp_r;
    }
    const int old_active = c->st.active;
    int new_active = old_active, rec_target = -1;
    time_slice_object(c, best.t);
    switch (best.kind) {
    case ECMC_EVENT_PAIR: {
        /* send_out_state, :154-190 */
        rec_target = best.target;
        c->stats.pair_events++;
        double dt = time_sub(best.t, in_stamp);
        for (int k = 0; k < npr; k++)
            for (int d = 0; d < D; d++) {
                double v = d == dir ? c->prog.speed : 0.0;
                in_state[k][d] = correct_position_entry(in_state[k][d] + v * dt, c->L);
            }
        double bounding_event_rate = 0.0, factor_derivative = 0.0;
        for (int i = 0; i < npr; i++)
            for (int j = 0; j < npr; j++) {
                int local = active_root * npr + i, target = best.target * npr + j;
                double sep[ECMC_MAX_DIM] = {0, 0, 0};
                separation_vector(in_state[i], c->pos + target * D, D, c->L, sep);
                double c1 = c->prog.pair_use_charge ? c->charge[local] : 1.0;
                double c2 = c->prog.pair_use_charge ? c->charge[target] : 1.0;
                double b = pot_derivative(&c->pair_bound, dir, c->prog.speed, sep, D, c1, c2);
                bounding_event_rate += b > 0.0 ? b : 0.0;
                factor_derivative += pot_derivative(&c->pair_pot, dir, c->prog.speed, sep, D, c1, c2);
            }
        if (factor_derivative > 0) {
            if (bounding_event_rate < factor_derivative) c->stats.bound_violations++;
            double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), 0);
            if (0 + (bounding_event_rate - 0) * u < factor_derivative) new_active = best.target * npr;
        }
        break;
    }
    case ECMC_EVENT_FACTOR_PAIR:
        /* RootUnitActiveTwoLeafUnitEventHandler.send_out_state (:102-125): the object of the target leaf takes over */
        rec_target = best.target;
        new_active = (best.target / npr) * npr;
        c->stats.factor_pair_events++;
        break;
    case ECMC_EVENT_END_OF_CHAIN:
        new_active = c->st.eoc_next_active;
        rec_target = new_active;
        c->stats.end_of_chain_events++;
        c->st.eoc_last_q = best.t.q; c->st.eoc_last_r = best.t.r;
        break;
    case ECMC_EVENT_SWITCH: {
        /* RootLeafUnitActiveSwitcher._send_out_state_leaf_unit_active (:129-169): random.choice over the leaves */
        uint32_t chosen = rng_randbelow(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_SWITCH, 0),
                                        (uint32_t)npr);
        new_active = active_root * npr + (int)chosen;
        c->st.mode = 0;
        break;
    }
    default: break;
    }
    if (rec) {
        memset(rec, 0, sizeof(*rec));
        rec->kind = best.kind;
        rec->target = rec_target;
        rec->target_cell = -1;
        rec->accepted = (best.kind == ECMC_EVENT_END_OF_CHAIN || best.kind == ECMC_EVENT_SWITCH) ? 1 : (new_active != old_active);
        rec->n_candidates = n_cand;
        rec->time_q = best.t.q; rec->time_r = best.t.r;
        for (int d = 0; d < D; d++) rec->active_pos[d] = c->pos[old_active * D + d];
    }
    c->st.event_counter++;
    c->stats.events++;
    c->stats.candidates += (uint64_t)n_cand;
    if (best.kind == ECMC_EVENT_END_OF_CHAIN) c->st.direction = (c->st.direction + 1) % D;
    molecule_occupancy_update(c, new_active);
    if (best.kind == ECMC_EVENT_SWITCH) {
        otime t = time_add(best.t, c->prog.switch_chain_length[0]);
        c->st.switch_q = t.q; c->st.switch_r = t.r;
    }
    if (best.kind == ECMC_EVENT_END_OF_CHAIN || best.kind == ECMC_EVENT_SWITCH) schedule_end_of_chain(c);
    if (rec) { rec->new_active = c->st.active; rec->new_direction = c->st.direction; rec->mode = c->st.mode; }
    return 1;
}


/* One iteration of SingleProcessMediator.run (single_process_mediator.py:91-156) restricted to device
 * events. Returns 0 if the next event time is not < until (nothing committed, candidate kept pending). */
/* ---- general velocities: two-dimensional hard disks tethered into dipoles, no cell system, the velocity rotated at
 * every end of chain (hard_disk_dipoles/hard_disk_dipoles.ini). Candidates of an event: the hard-sphere factor with
 * every leaf of every other object (factor type map entries between different objects, factor_type_maps.py:333-347,
 * recorded as ECMC_EVENT_FACTOR_PAIR), the tether with the partner leaf (ECMC_EVENT_BOND), the end of chain. Hard
 * potentials draw no random numbers (TwoLeafUnitEventHandler.send_event_time, two_leaf_unit_event_handler.py:105-138
 * with potential_change_required False) and always lift. */
static candidate disk_candidate(OrcChain *c, int kind, const opotential *pot, int target) {
    candidate cand;
    cand.kind = kind; cand.target = target; cand.target_cell = -1; cand.rate = 0.0;
    double sep[ECMC_MAX_DIM] = {0, 0, 0};
    separation_vector(c->pos + c->st.active * c->D, c->pos + target * c->D, c->D, c->L, sep);
    double velocity[ECMC_MAX_DIM] = {c->st.velocity[0], c->st.velocity[1], 0.0};
    double dt = pot->kind == ECMC_POT_HARD_SPHERE ? hs_displacement(pot->p0, velocity, sep, c->D)
                                                  : hd_displacement(pot->p0, pot->p1, velocity, sep, c->D);
    otime now = {c->st.time_q, c->st.time_r};
    cand.t = time_add(now, dt);
    return cand;
}

/* The velocity of a root unit after the velocity changes of an out-state, as _register_velocity_change_leaf_cnode and
 * _commit_sub_tree_non_leaf_velocity_change accumulate them (abstracts.py:165-227): `change` is the sum over the leaves
 * of (leaf change x leaf weight) in registration order; a unit without velocity takes the change, one with velocity adds
 * it and loses its velocity when every component is below 1e-13. Returns 0 if the unit ends without velocity. */
static int disk_root_velocity(double *velocity, int had_velocity, const double *change) {
    if (!had_velocity) { velocity[0] = change[0]; velocity[1] = change[1]; return 1; }
    velocity[0] += change[0]; velocity[1] += change[1];
    return !(fabs(velocity[0]) < 1.0e-13 && fabs(velocity[1]) < 1.0e-13);
}

static int disk_step(OrcChain *c, otime until, EcmcEventRecord *rec) {
    const int npr = c->npr, D = c->D;
    candidate best;
    int n_cand = 0;
    best.kind = ECMC_EVENT_NONE; best.t.q = ORC_INF; best.t.r = ORC_INF; best.target = -1; best.target_cell = -1;
    best.rate = 0.0;
    int was_pending = c->st.pending_kind != ECMC_EVENT_NONE;
    if (was_pending) {
        best.kind = c->st.pending_kind;
        best.t.q = c->st.pending_q; best.t.r = c->st.pending_r;
        best.target = c->st.pending_target;
    } else {
        const int active_root = c->st.active / npr, active_child = c->st.active % npr;
        /* factors between different objects: (child a, child b) entries with a = the active child */
        for (int root = 0; root < c->N / npr; root++) {
            if (root == active_root) continue;
            for (int f = 0; f < c->prog.n_inter_factors; f++) {
                if (c->prog.inter_factors[f][0] != active_child) continue;
                candidate cand = disk_candidate(c, ECMC_EVENT_FACTOR_PAIR, &c->inter_pot, root * npr + c->prog.inter_factors[f][1]);
                if (!isinf(cand.t.q)) { n_cand++; if (lt_candidate(&cand, &best)) best = cand; }
            }
        }
        for (int b = 0; b < c->prog.n_bonds; b++) {
            int partner = -1;
            if (c->prog.bonds[b][0] == active_child) partner = c->prog.bonds[b][1];
            else if (c->prog.bonds[b][1] == active_child) partner = c->prog.bonds[b][0];
            if (partner < 0) continue;
            candidate cand = disk_candidate(c, ECMC_EVENT_BOND, &c->bond_pot, active_root * npr + partner);
            if (!isinf(cand.t.q)) { n_cand++; if (lt_candidate(&cand, &best)) best = cand; }
        }
    }
    {
        candidate interaction = best;
        candidate cand;
        cand.kind = ECMC_EVENT_END_OF_CHAIN; cand.target = c->st.eoc_next_active; cand.target_cell = -1; cand.rate = 0.0;
        cand.t.q = c->st.eoc_q; cand.t.r = c->st.eoc_r;
        n_cand++;
        if (lt_candidate(&cand, &best)) best = cand;
        if (!time_lt(best.t, until)) {
            c->st.pending_kind = interaction.kind;
            c->st.pending_q = interaction.t.q; c->st.pending_r = interaction.t.r;
            c->st.pending_rate = 0.0;
            c->st.pending_target = interaction.target;
            if (!was_pending) {
                const double *pa = c->pos + c->st.active * D, *pr = c->root_pos + (c->st.active / npr) * D;
                c->st.pending_position = pa[0]; c->st.pending_position_y = pa[1];
                c->st.pending_root_position = pr[0]; c->st.pending_root_position_y = pr[1];
                c->st.pending_stamp_q = c->st.time_q;
                c->st.pending_stamp_r = c->st.time_r;
            }
            return 0;
        }
    }
    c->st.pending_kind = ECMC_EVENT_NONE;
    if (was_pending && best.kind != ECMC_EVENT_END_OF_CHAIN) {
        double *pa = c->pos + c->st.active * D, *pr = c->root_pos + (c->st.active / npr) * D;
        pa[0] = c->st.pending_position; pa[1] = c->st.pending_position_y;
        pr[0] = c->st.pending_root_position; pr[1] = c->st.pending_root_position_y;
        c->st.time_q = c->st.pending_stamp_q;
        c->st.time_r = c->st.pending_stamp_r;
    }
    const int old_active = c->st.active;
    int new_active = old_active;
    time_slice_active(c, best.t);
    const double old_velocity[2] = {c->st.velocity[0], c->st.velocity[1]};
    double new_velocity[2] = {old_velocity[0], old_velocity[1]};
    switch (best.kind) {
    case ECMC_EVENT_FACTOR_PAIR: new_active = best.target; c->stats.factor_pair_events++; break;
    case ECMC_EVENT_BOND: new_active = best.target; c->stats.bond_events++; break;
    case ECMC_EVENT_END_OF_CHAIN:
        /* _get_new_velocity, single_independent_active_sequential_direction_end_of_chain_event_handler.py:101-122 */
        new_velocity[0] = old_velocity[0] * c->prog.eoc_cos - old_velocity[1] * c->prog.eoc_sin;
        new_velocity[1] = old_velocity[0] * c->prog.eoc_sin + old_velocity[1] * c->prog.eoc_cos;
        new_active = c->st.eoc_next_active;
        c->stats.end_of_chain_events++;
        break;
    default: break;
    }
    {
        /* velocity changes of the leaves, -old for the old active leaf and +new for the new one (the same leaf at an end
         * of chain: -old + new, end_of_chain_event_handler.py:143-168; _exchange_velocity, abstracts.py:296-321), carried
         * to the root units with the leaf weight */
        const double weight = 1.0 / npr;
        const int old_root = old_active / npr, new_root = new_active / npr;
        double change[2];
        if (new_active == old_active) {
            if (best.kind == ECMC_EVENT_END_OF_CHAIN) {
                for (int d = 0; d < 2; d++) change[d] = (-old_velocity[d] + new_velocity[d]) * weight;
                disk_root_velocity(c->st.root_velocity, 1, change);
            }
        } else if (new_root == old_root) {
            for (int d = 0; d < 2; d++) {
                change[d] = -old_velocity[d] * weight;
                change[d] += new_velocity[d] * weight;
            }
            disk_root_velocity(c->st.root_velocity, 1, change);
        } else {
            /* the old root loses its velocity (|v w - v w| < 1e-13), the new root takes new x weight */
            for (int d = 0; d < 2; d++) c->st.root_velocity[d] = new_velocity[d] * weight;
        }
        c->st.velocity[0] = new_velocity[0]; c->st.velocity[1] = new_velocity[1];
    }
    if (rec) {
        memset(rec, 0, sizeof(*rec));
        rec->kind = best.kind;
        rec->target = best.kind == ECMC_EVENT_END_OF_CHAIN ? new_active : best.target;
        rec->target_cell = -1;
        rec->accepted = 1;
        rec->n_candidates = n_cand;
        rec->time_q = best.t.q; rec->time_r = best.t.r;
        for (int d = 0; d < D; d++) rec->active_pos[d] = c->pos[old_active * D + d];
    }
    c->st.event_counter++;
    c->stats.events++;
    c->stats.candidates += (uint64_t)n_cand;
    c->st.active = new_active;
    if (best.kind == ECMC_EVENT_END_OF_CHAIN) schedule_end_of_chain(c);
    if (rec) { rec->new_active = c->st.active; rec->new_direction = 0; }
    return 1;
}

static int chain_step(OrcChain *c, otime until, EcmcEventRecord *rec) {
    if (c->molecules) return molecule_step(c, until, rec);
    if (c->prog.eoc_sequential) return disk_step(c, until, rec);
    candidate best;
    int n_cand = 0;
    double boundary_position = 0.0;
    best.kind = ECMC_EVENT_NONE; best.t.q = ORC_INF; best.t.r = ORC_INF; best.target = -1; best.target_cell = -1;
    best.rate = 0.0;
    int was_pending = c->st.pending_kind != ECMC_EVENT_NONE;
    if (was_pending) {
        /* a candidate that survived a host control event: nothing is recomputed, no draws are consumed */
        best.kind = c->st.pending_kind;
        best.t.q = c->st.pending_q; best.t.r = c->st.pending_r;
        best.rate = c->st.pending_rate;
        if (best.kind == ECMC_EVENT_PAIR || best.kind == ECMC_EVENT_CELL_BOUNDING || best.kind == ECMC_EVENT_BOND)
            best.target = c->st.pending_target;
        else best.target_cell = c->st.pending_target;
        if (best.kind == ECMC_EVENT_CELL_BOUNDARY)
            boundary_position = c->cells.cell_min[best.target_cell * c->D + c->st.direction];
        n_cand = 0;
    } else {
        /* ExcludedCellsTagger, excluded_cells_tagger.py:129-132 */
        int nearby[125];
        int active_cell = c->st.active_cell;
        int nn = nearby_cells(&c->cells, active_cell, nearby);
        int m = c->prog.max_occupants;
        if (c->prog.pair_handler != ECMC_PAIR_NONE) {
            for (int i = 0; i < nn; i++)
                for (int s = 0; s < m; s++) {
                    int t = c->occ[nearby[i] * m + s];
                    if (t < 0) continue;
                    candidate cand;
                    cand.kind = ECMC_EVENT_PAIR; cand.target = t; cand.target_cell = -1; cand.rate = 0.0;
                    cand.t = pair_candidate_time(c, t);
                    if (!isinf(cand.t.q)) { /* heap_scheduler.py:139: only finite times are pushed */
                        n_cand++;
                        if (lt_candidate(&cand, &best)) best = cand;
                    }
                }
            /* SurplusCellsTagger, surplus_cells_tagger.py:129-131 */
            for (int s = 0; s < c->n_surplus; s++) {
                candidate cand;
                cand.kind = ECMC_EVENT_PAIR; cand.target = c->surplus[s]; cand.target_cell = -1; cand.rate = 0.0;
                cand.t = pair_candidate_time(c, cand.target);
                if (!isinf(cand.t.q)) {
                    n_cand++;
                    if (lt_candidate(&cand, &best)) best = cand;
                }
            }
        }
        /* factor-type-map factors that contain the active leaf */
        for (int b = 0; b < c->prog.n_bonds; b++) {
            int child = c->st.active % c->npr, root = c->st.active / c->npr;
            int partner = -1;
            if (c->prog.bonds[b][0] == child) partner = c->prog.bonds[b][1];
            else if (c->prog.bonds[b][1] == child) partner = c->prog.bonds[b][0];
            if (partner < 0) continue;
            candidate cand = bond_candidate(c, root * c->npr + partner);
            if (!isinf(cand.t.q)) { n_cand++; if (lt_candidate(&cand, &best)) best = cand; }
        }
        if (c->prog.veto_enabled == ECMC_FAR_CELL_BOUNDING) {
            /* CellBoundingPotentialTagger, cell_bounding_potential_tagger.py:150-155: every non-empty cell that is
             * not nearby the active cell */
            for (int cell = 0; cell < c->cells.n_cells; cell++) {
                int t = c->occ[cell * m];
                if (t < 0) continue;
                int is_near = 0;
                for (int i = 0; i < nn; i++) if (nearby[i] == cell) { is_near = 1; break; }
                if (is_near) continue;
                candidate cand = cell_bounding_candidate(c, t);
                if (!isinf(cand.t.q)) { n_cand++; if (lt_candidate(&cand, &best)) best = cand; }
            }
        } else if (c->prog.veto_enabled) {
            candidate cand = veto_candidate(c);
            if (!isinf(cand.t.q)) { n_cand++; if (lt_candidate(&cand, &best)) best = cand; }
        }
        if (!c->prog.no_cells) {
            candidate cand = boundary_candidate(c, &boundary_position);
            n_cand++;
            if (lt_candidate(&cand, &best)) best = cand;
        }
    }
    {
        /* the end-of-chain candidate lives in the scheduler since the chain started */
        candidate interaction = best;
        candidate cand;
        cand.kind = ECMC_EVENT_END_OF_CHAIN; cand.target = c->st.eoc_next_active; cand.target_cell = -1; cand.rate = 0.0;
        cand.t.q = c->st.eoc_q; cand.t.r = c->st.eoc_r;
        n_cand++;
        if (lt_candidate(&cand, &best)) best = cand;
        if (!time_lt(best.t, until)) {
            /* host control event first: the interaction winner stays in the scheduler (no trash), like the
             * end-of-chain candidate which persists anyway */
            c->st.pending_kind = interaction.kind;
            c->st.pending_q = interaction.t.q; c->st.pending_r = interaction.t.r;
            c->st.pending_rate = interaction.rate;
            c->st.pending_target = (interaction.kind == ECMC_EVENT_PAIR || interaction.kind == ECMC_EVENT_CELL_BOUNDING ||
                                    interaction.kind == ECMC_EVENT_BOND) ? interaction.target : interaction.target_cell;
            if (!was_pending) {
                /* the kept handlers hold copies of the in-state made before the control event time-slices
                 * the global state (single_process_mediator.py:105-109): their out-state starts from here */
                c->st.pending_position = c->pos[c->st.active * c->D + c->st.direction];
                if (c->npr > 1)
                    c->st.pending_root_position = c->root_pos[(c->st.active / c->npr) * c->D + c->st.direction];
                c->st.pending_stamp_q = c->st.time_q;
                c->st.pending_stamp_r = c->st.time_r;
            }
            return 0;
        }
    }
    c->st.pending_kind = ECMC_EVENT_NONE;
    if (was_pending && best.kind != ECMC_EVENT_END_OF_CHAIN) {
        /* the kept handler's stored in-state predates the control event's time slice; the end-of-chain
         * handler instead receives the current global state (mediator/mediator.py:233-249) */
        c->pos[c->st.active * c->D + c->st.direction] = c->st.pending_position;
        if (c->npr > 1) c->root_pos[(c->st.active / c->npr) * c->D + c->st.direction] = c->st.pending_root_position;
        c->st.time_q = c->st.pending_stamp_q;
        c->st.time_r = c->st.pending_stamp_r;
    }

    int old_active = c->st.active;
    int new_active = old_active;
    int accepted = 0;
    int rec_target = -1;
    time_slice_active(c, best.t);
    double c_act = c->charge[old_active];
    switch (best.kind) {
    case ECMC_EVENT_PAIR: {
        rec_target = best.target;
        if (c->prog.pair_handler == ECMC_PAIR_TWO_LEAF_UNIT) {
            /* TwoLeafUnitEventHandler.send_out_state, two_leaf_unit_event_handler.py:140-154 */
            accepted = 1;
        } else {
            /* TwoLeafUnitBoundingPotentialEventHandler.send_out_state (:148-168) +
             * _calculate_out_state_of_two_leaf_unit_bounding_potential (event_handler_with_bounding_potential.py:75-101) */
            double sep[ECMC_MAX_DIM] = {0, 0, 0};
            separation_vector(c->pos + old_active * c->D, c->pos + best.target * c->D, c->D, c->L, sep);
            double c1 = c->prog.pair_use_charge ? c_act : 1.0;
            double c2 = c->prog.pair_use_charge ? c->charge[best.target] : 1.0;
            double bounding_rate = pot_derivative(&c->pair_bound, c->st.direction, c->prog.speed, sep, c->D, c1, c2);
            double real = pot_derivative(&c->pair_pot, c->st.direction, c->prog.speed, sep, c->D, c1, c2);
            if (real > 0) {
                if (bounding_rate < real) c->stats.bound_violations++;
                double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), 0);
                if (0 + (bounding_rate - 0) * u < real) accepted = 1;
            }
        }
        if (accepted) new_active = best.target;
        c->stats.pair_events++;
        break;
    }
    case ECMC_EVENT_CELL_VETO: {
        /* mediator.get_arguments_cell_veto_event_handler (mediator/mediator.py:265-292) +
         * LeafUnitCellVetoEventHandler.send_out_state (leaf_unit_cell_veto_event_handler.py:117-149) */
        int t = c->occ[best.target_cell * c->prog.max_occupants];
        rec_target = t;
        if (t >= 0) {
            double sep[ECMC_MAX_DIM] = {0, 0, 0};
            separation_vector(c->pos + old_active * c->D, c->pos + t * c->D, c->D, c->L, sep);
            double c1 = c->prog.veto_use_charge ? c_act : 1.0;
            double c2 = c->prog.veto_use_charge ? c->charge[t] : 1.0;
            double real = pot_derivative(&c->veto_pot, c->st.direction, c->prog.speed, sep, c->D, c1, c2);
            if (real > 0) {
                if (best.rate < real) c->stats.bound_violations++;
                double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), 0);
                if (0 + (best.rate - 0) * u < real) { accepted = 1; new_active = t; }
            }
        }
        c->stats.veto_events++;
        if (accepted) c->stats.veto_accepted++;
        break;
    }
    case ECMC_EVENT_CELL_BOUNDING: {
        /* TwoLeafUnitCellBoundingPotentialEventHandler.send_out_state (:179-211): the bounding event rate is the one
         * stored by send_event_time (times the speed, StandardVelocityPotential.derivative), confirmed against the
         * real potential like any bounded pair (event_handler_with_bounding_potential.py:75-101) */
        rec_target = best.target;
        double sep[ECMC_MAX_DIM] = {0, 0, 0};
        separation_vector(c->pos + old_active * c->D, c->pos + best.target * c->D, c->D, c->L, sep);
        double c1 = c->prog.veto_use_charge ? c_act : 1.0;
        double c2 = c->prog.veto_use_charge ? c->charge[best.target] : 1.0;
        double bounding_rate = best.rate * c->prog.speed;
        double real = pot_derivative(&c->veto_pot, c->st.direction, c->prog.speed, sep, c->D, c1, c2);
        if (real > 0) {
            if (bounding_rate < real) c->stats.bound_violations++;
            double u = rng_double(c->prog.seed, c->st.stream, c->st.event_counter, ECMC_SLOT(ECMC_SLOT_CONFIRM, 0), 0);
            if (0 + (bounding_rate - 0) * u < real) accepted = 1;
        }
        if (accepted) new_active = best.target;
        c->stats.pair_events++;
        break;
    }
    case ECMC_EVENT_BOND:
        /* TwoLeafUnitEventHandler.send_out_state, two_leaf_unit_event_handler.py:140-154 */
        rec_target = best.target;
        accepted = 1;
        new_active = best.target;
        c->stats.bond_events++;
        break;
    case ECMC_EVENT_CELL_BOUNDARY:
        /* CellBoundaryEventHandler.send_out_state, cell_boundary_event_handler.py:158-173 */
        c->pos[old_active * c->D + c->st.direction] = boundary_position;
        c->stats.boundary_events++;
        break;
    case ECMC_EVENT_END_OF_CHAIN:
        /* EndOfChainEventHandler.send_out_state (abstracts/end_of_chain_event_handler.py:107-187) with
         * _get_new_velocity (single_independent_active_periodic_direction_...:183-201) */
        new_active = c->st.eoc_next_active;
        rec_target = new_active;
        accepted = 1;
        c->stats.end_of_chain_events++;
        break;
    default: break;
    }
    if (rec) {
        memset(rec, 0, sizeof(*rec));
        rec->kind = best.kind;
        rec->target = rec_target;
        rec->target_cell = best.target_cell;
        rec->accepted = accepted;
        rec->n_candidates = n_cand;
        rec->time_q = best.t.q; rec->time_r = best.t.r;
        for (int d = 0; d < c->D; d++) rec->active_pos[d] = c->pos[old_active * c->D + d];
    }
    c->st.event_counter++;
    c->stats.events++;
    c->stats.candidates += (uint64_t)n_cand;
    if (best.kind == ECMC_EVENT_END_OF_CHAIN) c->st.direction = (c->st.direction + 1) % c->D;
    occupancy_update(c, new_active);
    if (best.kind == ECMC_EVENT_END_OF_CHAIN) schedule_end_of_chain(c);
    if (rec) { rec->new_active = c->st.active; rec->new_direction = c->st.direction; }
    return 1;
}

/* Advance until the next event time reaches `until` or max_events were committed. Returns #events. */
ORC_API int64_t orc_chain_run(OrcChain *c, double until_q, double until_r, int64_t max_events, EcmcEventRecord *records,
                              int64_t max_records) {
    otime until = {until_q, until_r};
    int64_t n = 0;
    while (max_events <= 0 || n < max_events) {
        EcmcEventRecord *rec = (records && n < max_records) ? records + n : NULL;
        if (!chain_step(c, until, rec)) {
            /* sampling / end-of-run handler time-slices the active unit
             * (fixed_interval_sampling_event_handler.py:96-109) */
            time_slice_active(c, until);
            break;
        }
        n++;
    }
    return n;
}
