/*
 * mate_oracle.c -- CPU restatement (float64, scalar C) of the reference's
 * MultiAgentTracking step path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call this file.  The product (mate_b200/) never does: it
 * fails loudly when its CUDA library is missing.
 *
 * PARITY PINNING: the reference (XuehaiPan/mate) ships no tests, golden vectors or
 * known-answer values for this path, so the oracle is pinned against outputs of the
 * reference itself, produced by oracle/gen_golden.py running the UNMODIFIED reference in
 * the build container and committed under tests/golden/ (tests/test_oracle_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root, e.g. mate/entities.py).  The arithmetic deliberately mirrors the
 * reference's float64 operation order, including the lazy polar/cartesian caching of
 * mate.utils.Vector2D (mate/utils.py:161-271); it is written for fidelity, not speed.
 * Field-of-view occlusion is served exactly like the reference: a per-camera sampled
 * (phi, rho) polyline built at reset (mate/entities.py:362-479) and queried with
 * np.interp semantics (mate/entities.py:507-511).
 *
 * The reset path restates the reference's reset ALGORITHM (mate/environment.py:679-834)
 * but draws from the counter-based Philox4x32-10 streams the CUDA path uses (the
 * reference's MT19937 streams behind gym.spaces.Box.sample are not reproducible on a
 * GPU); it is therefore an oracle for the CUDA reset and pinned to the reference only
 * distributionally (tests/test_reset_distribution.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mate_b200.h"

#define NW MATE_NUM_WAREHOUSES
#define MAXC MATE_MAX_CAMERAS
#define MAXT MATE_MAX_TARGETS
#define MAXO MATE_MAX_OBSTACLES
#define MAX_RAYS_PER_OBSTACLE 186 /* 4 edge rays + at most max(16, int(2*half))+1 <= 181 lattice rays */
#define NUM_RESET_RETRIES 500     /* mate/environment.py:53 */

static const double PI = 3.14159265358979323846;
#define RAD2DEG (180.0 / PI) /* mate/utils.py:63 */
#define DEG2RAD (PI / 180.0) /* mate/utils.py:68 */

static const double TERRAIN_SIZE = 1000.0;     /* mate/constants.py:52 */
static const double WAREHOUSE_RADIUS = 75.0;   /* mate/constants.py:67 */
static const double WAREHOUSES[NW][2] = {      /* mate/constants.py:70-72 */
    {925.0, 925.0}, {-925.0, 925.0}, {-925.0, -925.0}, {925.0, -925.0}};

typedef struct {
    double cam_x[MAXC], cam_y[MAXC], cam_phi[MAXC], cam_theta[MAXC], cam_rs[MAXC];
    double tgt_x[MAXT], tgt_y[MAXT], tgt_orient[MAXT];
    int tgt_cap[MAXT], tgt_goal[MAXT], tgt_weight[MAXT], tgt_bounty[MAXT], tgt_empty[MAXT];
    int tgt_colliding[MAXT], tgt_done[MAXT];
    double obs_x[MAXO], obs_y[MAXO], obs_r[MAXO];
    int remaining[NW][NW], awaiting[NW];
    int delivered, episode_step, episode_id;
    double ep_reward, delayed_ep_reward;
    /* derived by update_view / assign_goals / observe */
    uint8_t m_ct[MAXC][MAXT], m_cc[MAXC][MAXC], m_co[MAXC][MAXO];
    uint8_t m_tc[MAXT][MAXC], m_to[MAXT][MAXO], m_tt[MAXT][MAXT];
    uint8_t tracked[MAXT];
    double wh_dist[MAXT][NW];
    double coverage, real_coverage, transport;
    double coverage_sum; /* running sum of coverage_rate over the episode (stats) */
    /* FOV polyline per camera */
    int fov_n[MAXC];
    double* fov_phi[MAXC];
    double* fov_rho[MAXC];
} Env;

typedef struct {
    MateConfig cfg;
    double cam_ranges[MAXC][4], tgt_ranges[MAXT][4], obs_ranges[MAXO][4];
    int num_envs;
    int64_t env_index_base;
    int dc, dt;
    double freight_scale, bounty_scale, reward_scale;
    int fov_cap;
    Env* envs;
    double stats[16];
    int num_threads;
} Oracle;

/* ------------------------------------------------------------------ math helpers */

/* Python float modulo: result takes the sign of the divisor (used by normalize_angle). */
static double pymod(double x, double y) {
    double m = fmod(x, y);
    if (m != 0.0) {
        if ((y < 0.0) != (m < 0.0)) m += y;
    } else {
        m = copysign(0.0, y);
    }
    return m;
}

/* mate/utils.py:155-158 */
static double normalize_angle(double a) { return pymod(a + 180.0, 360.0) - 180.0; }

/* np.linalg.norm of a 2-vector (mate/entities.py:91-94, mate/utils.py:217-221) */
static double norm2(double x, double y) { return sqrt(x * x + y * y); }

/* mate/utils.py:124-131 */
static double arctan2_deg(double y, double x) { return atan2(y, x) * RAD2DEG; }

/* mate/utils.py:144-152 */
static void polar2cartesian(double rho, double phi, double* x, double* y) {
    double r = phi * DEG2RAD;
    *x = rho * cos(r);
    *y = rho * sin(r);
}

/* ------------------------------------------------------------------ Philox4x32-10 */
/* Counter-based RNG (Salmon et al., SC'11).  Shared draw scheme with the CUDA path:
 *   key = (seed_lo, seed_hi); counter = (index, stream, global_env_index, episode_id). */
static void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                       uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum {
    STREAM_SHUFFLE_CAM = 0, STREAM_SHUFFLE_TGT = 1, STREAM_SHUFFLE_OBS = 2, STREAM_CAPACITY = 3,
    STREAM_PLACE = 4, STREAM_CARGO = 5, STREAM_INIT_GOAL = 6, STREAM_TRANSMIT = 7, STREAM_CHOICE = 8
};

typedef struct { uint64_t seed; uint32_t env, episode; } RngKey;

static void rng_words(const RngKey* k, uint32_t stream, uint32_t index, uint32_t out[4]) {
    philox4x32(index, stream, k->env, k->episode, (uint32_t)k->seed, (uint32_t)(k->seed >> 32), out);
}
/* uniform double in [0, 1) with 53 random bits */
static double rng_u01(const RngKey* k, uint32_t stream, uint32_t index) {
    uint32_t w[4];
    rng_words(k, stream, index, w);
    uint64_t bits = ((uint64_t)w[1] << 32) | w[0];
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}
/* integer in [0, n) */
static uint32_t rng_below(const RngKey* k, uint32_t stream, uint32_t index, uint32_t n) {
    uint32_t w[4];
    rng_words(k, stream, index, w);
    return (uint32_t)(((uint64_t)w[0] * n) >> 32);
}

/* ------------------------------------------------------------------ Obstacle.obstruct */

/* Lazy-cached 2-D vector, mirroring mate.utils.Vector2D (mate/utils.py:161-271). */
typedef struct {
    double vx, vy;   /* _vector (valid if has_v)  */
    double norm;     /* _norm   (valid if has_n)  */
    double angle;    /* _angle  (valid if has_a)  */
    int has_v, has_n, has_a;
} Vec;

static void vec_from_vector(Vec* v, double x, double y) {
    v->vx = x; v->vy = y; v->has_v = 1; v->has_n = 0; v->has_a = 0; v->norm = 0; v->angle = 0;
}
static void vec_need_v(Vec* v) { /* mate/utils.py:177-181 */
    if (!v->has_v) { polar2cartesian(v->norm, v->angle, &v->vx, &v->vy); v->has_v = 1; }
}
static double vec_norm(Vec* v) { /* mate/utils.py:217-221 */
    if (!v->has_n) { v->norm = norm2(v->vx, v->vy); v->has_n = 1; }
    return v->norm;
}
static double vec_angle(Vec* v) { /* mate/utils.py:206-210 */
    if (!v->has_a) { v->angle = arctan2_deg(v->vy, v->vx); v->has_a = 1; }
    return v->angle;
}
static void vec_set_norm(Vec* v, double value) { /* mate/utils.py:223-229 */
    double angle = vec_angle(v);
    v->norm = fabs(value); v->has_n = 1;
    v->has_v = 0;
    if (value < 0.0) { v->angle = normalize_angle(angle + 180.0); v->has_a = 1; }
}

/* mate/entities.py:158-184.  `ray` starts at (ox, oy); disc centre (px, py), radius R. */
static void obstruct(Vec* ray, double ox, double oy, double px, double py, double R,
                     int keep_tangential) {
    double relx = px - ox, rely = py - oy;
    double reln = norm2(relx, rely);
    double norm = vec_norm(ray);
    if (norm == 0.0 || reln < R) { /* return -ray */
        vec_need_v(ray);
        vec_from_vector(ray, -ray->vx, -ray->vy);
        return;
    }
    if (reln >= norm + R) return;
    vec_need_v(ray);
    double inner = relx * ray->vx + rely * ray->vy;
    if (inner >= 0.0) {
        double c = fmin(1.0, inner / (reln * norm));
        double perpendicular = reln * sqrt(1.0 - c * c);
        if (R > perpendicular) {
            double half_chord = sqrt(R * R - perpendicular * perpendicular);
            double new_norm = fmax(0.0, reln * c - half_chord);
            if (new_norm < norm) {
                double oldx = ray->vx, oldy = ray->vy;
                vec_set_norm(ray, new_norm);
                if (keep_tangential) {
                    vec_need_v(ray);
                    double radx = (ox + ray->vx) - px, rady = (oy + ray->vy) - py;
                    double k = (norm - new_norm) * half_chord / (R * R);
                    vec_from_vector(ray, oldx + radx * k, oldy + rady * k);
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ FOV polyline (A7) */

typedef struct { double angle, norm; int tangent_of; } Ray; /* tangent_of: obstacle whose exact tangent ray this is, else -1 */

static int ray_cmp(const void* a, const void* b) {
    const Ray* x = (const Ray*)a; const Ray* y = (const Ray*)b;
    if (x->angle < y->angle) return -1;
    if (x->angle > y->angle) return 1;
    if (x->norm < y->norm) return -1; /* equal angles: smaller norm first (dedupe keeps the min) */
    if (x->norm > y->norm) return 1;
    return 0;
}

/* Camera.reset default boundary + Camera.add_obstacles (mate/entities.py:336-345, 362-479). */
static void build_fov(const Oracle* S, Env* e, int c) {
    const MateConfig* cfg = &S->cfg;
    const double Rmax = cfg->camera_max_sight_range;
    const double cx = e->cam_x[c], cy = e->cam_y[c];
    Ray* rays = (Ray*)malloc(sizeof(Ray) * (size_t)S->fov_cap);
    int n = 0, inside = 0;
    int vis[MAXO], nvis = 0;
    for (int k = 0; k < 360; ++k) { /* entities.py:336-339: linspace(-180, 180, 360, endpoint=False) */
        rays[n].angle = normalize_angle(-180.0 + (double)k * 1.0);
        rays[n].norm = Rmax;
        rays[n].tangent_of = -1;
        ++n;
    }
    for (int o = 0; o < cfg->num_obstacles; ++o) { /* entities.py:363-368 (strict <) */
        double d = norm2(cx - e->obs_x[o], cy - e->obs_y[o]);
        int visible = d < Rmax + e->obs_r[o];
        e->m_co[c][o] = (uint8_t)visible;
        if (visible) vis[nvis++] = o;
    }
    if (cfg->obstacle_transmittance != 1.0) {
        for (int i = 0; i < nvis && !inside; ++i) { /* entities.py:373-417 */
            int o = vis[i];
            double relx = e->obs_x[o] - cx, rely = e->obs_y[o] - cy;
            double reln = norm2(relx, rely);
            double R = e->obs_r[o];
            if (R > reln) { inside = 1; break; } /* entities.py:378-387 */
            double half = asin(R / reln) * RAD2DEG;
            double max_rho = fmin(Rmax, reln + R);
            double rel_angle = arctan2_deg(rely, relx);
            double left = rel_angle - half, right = rel_angle + half;
            double edges[4] = {left - 0.01, left + 0.01, right - 0.01, right + 0.01};
            for (int k = 0; k < 4; ++k) {
                rays[n].angle = normalize_angle(edges[k]); rays[n].norm = Rmax; rays[n].tangent_of = -1; ++n;
            }
            int num = ((int)(2.0 * half) > 16 ? (int)(2.0 * half) : 16) + 1;
            /* np.linspace(left, right, num): y = arange(num) * step + start; y[-1] = stop */
            double step = (right - left) / (double)(num - 1);
            for (int k = 0; k < num; ++k) {
                double a = (k == num - 1) ? right : ((double)k * step + left);
                if (step == 0.0 && k != num - 1) a = ((double)k / (double)(num - 1)) * (right - left) + left;
                rays[n].angle = normalize_angle(a); rays[n].norm = max_rho;
                rays[n].tangent_of = (k == 0 || k == num - 1) ? o : -1; ++n;
            }
        }
        if (inside) { /* camera inside an obstacle: sight range 0 everywhere */
            n = 0;
            rays[n].angle = -180.0; rays[n].norm = 0.0; rays[n].tangent_of = -1; ++n;
        } else {
            for (int i = 0; i < nvis; ++i) { /* entities.py:450-455: obstruct every ray */
                int o = vis[i];
                for (int k = 0; k < n; ++k) {
                    /* DELIBERATE DEVIATION (documented in DESIGN.md "tangent rays"): the two end
                     * rays of an obstacle's linspace lattice are exactly tangent to that obstacle
                     * (perpendicular == radius mathematically), so in the reference the test
                     * `self.radius > perpendicular` (entities.py:170) is decided by the last-bit
                     * rounding of numpy's sin/cos and is not reproducible across platforms.  The
                     * oracle and the CUDA path both fix the exact-arithmetic outcome: a tangent
                     * ray is not shortened by its own obstacle. */
                    if (rays[k].tangent_of == o) continue;
                    Vec v;
                    v.has_v = 0; v.angle = rays[k].angle; v.has_a = 1; v.norm = rays[k].norm; v.has_n = 1;
                    v.vx = v.vy = 0.0;
                    obstruct(&v, cx, cy, e->obs_x[o], e->obs_y[o], e->obs_r[o], 0);
                    /* camera not inside the disc and norm > 0 here, so the angle is unchanged;
                       a zero-length ray would be "reversed" (-0 vector): keep norm 0. */
                    rays[k].norm = v.has_n ? v.norm : 0.0;
                }
            }
        }
    }
    qsort(rays, (size_t)n, sizeof(Ray), ray_cmp); /* entities.py:458 */
    int m = 0;
    for (int k = 0; k < n; ++k) { /* entities.py:460-466: equal angles keep the smaller norm */
        if (m > 0 && e->fov_phi[c][m - 1] == rays[k].angle) {
            if (e->fov_rho[c][m - 1] > rays[k].norm) e->fov_rho[c][m - 1] = rays[k].norm;
        } else {
            e->fov_phi[c][m] = rays[k].angle; e->fov_rho[c][m] = rays[k].norm; ++m;
        }
    }
    e->fov_phi[c][m] = e->fov_phi[c][0] + 360.0; /* entities.py:470-471 */
    e->fov_rho[c][m] = e->fov_rho[c][0];
    e->fov_n[c] = m + 1;
    free(rays);
}

/* np.interp (numpy/_core/src/multiarray/compiled_base.c arr_interp), which
 * scipy.interpolate.interp1d(kind='linear') delegates to (mate/entities.py:476, 507-511). */
static double np_interp(double x, const double* xp, const double* fp, int n) {
    if (x <= xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    int lo = 0, hi = n - 1; /* invariant: xp[lo] <= x < xp[hi] */
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    if (xp[lo] == x) return fp[lo];
    double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
    return slope * (x - xp[lo]) + fp[lo];
}

/* ------------------------------------------------------------------ step path */

/* Camera.simulate (mate/entities.py:347-360) */
static void camera_simulate(const MateConfig* cfg, Env* e, int c, double a0, double a1) {
    double da = fmin(fmax(a0, -cfg->camera_rotation_step), cfg->camera_rotation_step);
    double dv = fmin(fmax(a1, -cfg->camera_zooming_step), cfg->camera_zooming_step);
    e->cam_phi[c] = normalize_angle(e->cam_phi[c] + da);
    e->cam_theta[c] = fmin(fmax(e->cam_theta[c] + dv, cfg->camera_min_viewing_angle), 180.0);
    double area_product = cfg->camera_min_viewing_angle * (cfg->camera_max_sight_range * cfg->camera_max_sight_range);
    e->cam_rs[c] = sqrt(area_product / e->cam_theta[c]);
}

/* Target.simulate (mate/entities.py:645-668); brute force over all discs (obstacles then
 * camera barriers, mate/environment.py:743) instead of the spatial hash broad phase. */
static void target_simulate(const MateConfig* cfg, Env* e, int t, double ax, double ay) {
    double step_size = cfg->target_step_size / (double)e->tgt_cap[t]; /* entities.py:612-615 */
    double ox = e->tgt_x[t], oy = e->tgt_y[t];
    Vec step;
    vec_from_vector(&step, ax, ay);
    if (vec_norm(&step) > step_size) vec_set_norm(&step, step_size);
    vec_need_v(&step);
    double desx = ox + step.vx, desy = oy + step.vy;
    for (int o = 0; o < cfg->num_obstacles; ++o)
        obstruct(&step, ox, oy, e->obs_x[o], e->obs_y[o], e->obs_r[o], 1);
    for (int c = 0; c < cfg->num_cameras; ++c)
        obstruct(&step, ox, oy, e->cam_x[c], e->cam_y[c], cfg->camera_radius, 1);
    vec_need_v(&step);
    double nx = fmin(fmax(ox + step.vx, -TERRAIN_SIZE), TERRAIN_SIZE);
    double ny = fmin(fmax(oy + step.vy, -TERRAIN_SIZE), TERRAIN_SIZE);
    e->tgt_x[t] = nx; e->tgt_y[t] = ny;
    e->tgt_colliding[t] = !(fabs(nx - desx) <= 1e-6 && fabs(ny - desy) <= 1e-6);
    if (nx != ox || ny != oy) /* mate/environment.py:1349-1352 */
        e->tgt_orient[t] = arctan2_deg(ny - oy, nx - ox);
}

/* Camera.perceive (mate/entities.py:491-505).  draw: -1 => evaluate `transmit` lazily */
static int camera_perceive(const Oracle* S, const Env* e, int c, double qx, double qy,
                           int transmit_outcome) {
    double relx = qx - e->cam_x[c], rely = qy - e->cam_y[c];
    double dist = norm2(relx, rely);
    if (dist > e->cam_rs[c]) return 0;
    double ang = arctan2_deg(rely, relx);
    double ra = fabs(e->cam_phi[c] - ang);
    ra = fmin(ra, 360.0 - ra);
    if (ra * 2.0 > e->cam_theta[c]) return 0;
    if (transmit_outcome != 0) return 1;
    (void)S;
    double range = np_interp(normalize_angle(ang), e->fov_phi[c], e->fov_rho[c], e->fov_n[c]);
    return dist <= range * (1.0 + 1e-6);
}

/* reached the binomial draw? (first two tests of Camera.perceive) */
static int camera_reaches_draw(const Env* e, int c, double qx, double qy) {
    double relx = qx - e->cam_x[c], rely = qy - e->cam_y[c];
    double dist = norm2(relx, rely);
    if (dist > e->cam_rs[c]) return 0;
    double ang = arctan2_deg(rely, relx);
    double ra = fabs(e->cam_phi[c] - ang);
    ra = fmin(ra, 360.0 - ra);
    return !(ra * 2.0 > e->cam_theta[c]);
}

/* _update_view (mate/environment.py:1356-1388) */
static void update_view(const Oracle* S, Env* e, const uint8_t* transmit, const RngKey* key, int draw_step) {
    const MateConfig* cfg = &S->cfg;
    const int nc = cfg->num_cameras, nt = cfg->num_targets, no = cfg->num_obstacles;
    for (int t = 0; t < nt; ++t) {
        for (int c = 0; c < nc; ++c) {
            int outcome = 0;
            if (camera_reaches_draw(e, c, e->tgt_x[t], e->tgt_y[t])) {
                if (transmit) {
                    outcome = transmit[c * nt + t];
                } else { /* binomial(1, p): u < p */
                    uint32_t index = (uint32_t)draw_step * (uint32_t)(nc * nt) + (uint32_t)(c * nt + t);
                    outcome = rng_u01(key, STREAM_TRANSMIT, index) < cfg->obstacle_transmittance;
                }
            }
            e->m_ct[c][t] = (uint8_t)camera_perceive(S, e, c, e->tgt_x[t], e->tgt_y[t], outcome);
            /* Sensor.perceive (mate/entities.py:229-232) */
            e->m_tc[t][c] = norm2(e->tgt_x[t] - e->cam_x[c], e->tgt_y[t] - e->cam_y[c]) <=
                            cfg->target_sight_range + cfg->camera_radius;
        }
        for (int o = 0; o < no; ++o)
            e->m_to[t][o] = norm2(e->tgt_x[t] - e->obs_x[o], e->tgt_y[t] - e->obs_y[o]) <=
                            cfg->target_sight_range + e->obs_r[o];
        for (int u = 0; u < nt; ++u)
            e->m_tt[t][u] = (t == u) || norm2(e->tgt_x[t] - e->tgt_x[u], e->tgt_y[t] - e->tgt_y[u]) <=
                                            cfg->target_sight_range + 0.0;
    }
    for (int c = 0; c < nc; ++c)
        for (int d = 0; d < nc; ++d) /* camera->camera uses transmittance 0.0: binomial == 0 */
            e->m_cc[c][d] = (c == d) || camera_perceive(S, e, c, e->cam_x[d], e->cam_y[d], 0);
    for (int t = 0; t < nt; ++t) {
        int any = 0;
        for (int c = 0; c < nc; ++c) any |= e->m_ct[c][t];
        e->tracked[t] = (uint8_t)any;
    }
}

/* _assign_goals (mate/environment.py:1271-1324) */
static void assign_goals(const Oracle* S, Env* e, const int8_t* goal_choice, const RngKey* key,
                         int draw_step, double* reward, double* delayed_reward) {
    const MateConfig* cfg = &S->cfg;
    const int nt = cfg->num_targets;
    int old_goals[MAXT];
    double delayed = 0.0, r = 0.0;
    for (int t = 0; t < nt; ++t) {
        old_goals[t] = e->tgt_goal[t];
        if (e->tracked[t] && e->tgt_bounty[t] > 0) r -= 1.0;
    }
    for (int t = 0; t < nt; ++t) {
        int b = e->tgt_bounty[t] - (int)e->tracked[t];
        e->tgt_bounty[t] = b > 0 ? b : 0;
    }
    for (int t = 0; t < nt; ++t) {
        int goal = old_goals[t]; /* zip() captured the pre-loop value (environment.py:1278-1280) */
        int capacity = e->tgt_cap[t];
        int inside[NW];
        for (int w = 0; w < NW; ++w) {
            double dx = e->tgt_x[t] - WAREHOUSES[w][0], dy = e->tgt_y[t] - WAREHOUSES[w][1];
            e->wh_dist[t][w] = norm2(dx, dy);
            inside[w] = fmax(fabs(dx), fabs(dy)) <= WAREHOUSE_RADIUS;
        }
        for (int w = 0; w < NW; ++w) {
            if (!inside[w]) continue;
            if (goal >= 0) {
                if (goal == w) {
                    int weight = e->tgt_weight[t];
                    double total_bounty = weight * S->bounty_scale;
                    double freight = weight * S->freight_scale; /* freights[t], environment.py:1310 */
                    double rew = freight + e->tgt_bounty[t];
                    r += rew;
                    delayed += rew - (total_bounty - e->tgt_bounty[t]);
                    e->delivered += weight;
                    e->awaiting[goal] -= weight;
                } else {
                    continue;
                }
            }
            e->tgt_bounty[t] = 0; e->tgt_weight[t] = 0; e->tgt_goal[t] = -1;
            int any = 0;
            for (int g = 0; g < NW; ++g) any |= e->remaining[w][g] > 0;
            if (any) {
                int new_goal;
                if (goal_choice) {
                    new_goal = goal_choice[t];
                } else { /* np_random.choice(flatnonzero(remaining[w] > 0)) */
                    int cand[NW], ncand = 0;
                    for (int g = 0; g < NW; ++g) if (e->remaining[w][g] > 0) cand[ncand++] = g;
                    uint32_t index = (uint32_t)draw_step * (uint32_t)nt + (uint32_t)t;
                    new_goal = cand[rng_below(key, STREAM_CHOICE, index, (uint32_t)ncand)];
                }
                int rem = e->remaining[w][new_goal];
                int weight = capacity < rem ? capacity : rem;
                e->remaining[w][new_goal] -= weight;
                e->tgt_weight[t] = weight;
                e->tgt_bounty[t] = (int)(weight * S->bounty_scale);
                e->tgt_goal[t] = new_goal;
                break;
            }
        }
        for (int w = 0; w < NW; ++w) { /* environment.py:1317-1318 */
            if (!inside[w]) continue;
            int any = 0;
            for (int g = 0; g < NW; ++g) any |= e->remaining[w][g] > 0;
            if (any) e->tgt_empty[t] &= ~(1 << w); else e->tgt_empty[t] |= (1 << w);
        }
    }
    for (int t = 0; t < nt; ++t)
        e->tgt_done[t] = (e->tgt_goal[t] != old_goals[t]) && (old_goals[t] >= 0);
    *reward = r;
    *delayed_reward = delayed;
}

/* joint_observation (mate/environment.py:908-983) with entity state()s
 * (mate/entities.py:313-324 camera, :631-637 target, :147-148 obstacle). */
static void joint_observation(const Oracle* S, Env* e, double* cam_obs, double* tgt_obs) {
    const MateConfig* cfg = &S->cfg;
    const int nc = cfg->num_cameras, nt = cfg->num_targets, no = cfg->num_obstacles;
    double preserved[13];
    preserved[0] = nc; preserved[1] = nt; preserved[2] = no; preserved[3] = 0.0;
    for (int w = 0; w < NW; ++w) { preserved[4 + 2 * w] = WAREHOUSES[w][0]; preserved[5 + 2 * w] = WAREHOUSES[w][1]; }
    preserved[12] = WAREHOUSE_RADIUS;
    double cam_pub[MAXC][7], tgt_pub[MAXT][5];
    for (int c = 0; c < nc; ++c) {
        double sx, sy;
        polar2cartesian(e->cam_rs[c], e->cam_phi[c], &sx, &sy);
        cam_pub[c][0] = e->cam_x[c]; cam_pub[c][1] = e->cam_y[c]; cam_pub[c][2] = cfg->camera_radius;
        cam_pub[c][3] = sx; cam_pub[c][4] = sy; cam_pub[c][5] = e->cam_theta[c]; cam_pub[c][6] = 1.0;
    }
    for (int t = 0; t < nt; ++t) {
        tgt_pub[t][0] = e->tgt_x[t]; tgt_pub[t][1] = e->tgt_y[t]; tgt_pub[t][2] = cfg->target_sight_range;
        tgt_pub[t][3] = (e->tgt_goal[t] >= 0 && e->tgt_weight[t] > 0) ? 1.0 : 0.0; tgt_pub[t][4] = 1.0;
    }
    if (cam_obs) {
        for (int c = 0; c < nc; ++c) {
            double* row = cam_obs + (size_t)c * S->dc;
            int k = 0;
            for (int i = 0; i < 13; ++i) row[k++] = preserved[i];
            row[3] = (double)c;
            for (int i = 0; i < 6; ++i) row[k++] = cam_pub[c][i];
            row[k++] = cfg->camera_max_sight_range; row[k++] = cfg->camera_rotation_step; row[k++] = cfg->camera_zooming_step;
            for (int t = 0; t < nt; ++t) for (int i = 0; i < 5; ++i) row[k++] = e->m_ct[c][t] ? tgt_pub[t][i] : 0.0;
            for (int o = 0; o < no; ++o) {
                int m = e->m_co[c][o];
                row[k++] = m ? e->obs_x[o] : 0.0; row[k++] = m ? e->obs_y[o] : 0.0;
                row[k++] = m ? e->obs_r[o] : 0.0; row[k++] = m ? 1.0 : 0.0;
            }
            for (int d = 0; d < nc; ++d) for (int i = 0; i < 7; ++i) row[k++] = e->m_cc[c][d] ? cam_pub[d][i] : 0.0;
        }
    }
    if (tgt_obs) {
        for (int t = 0; t < nt; ++t) {
            double* row = tgt_obs + (size_t)t * S->dt;
            int k = 0;
            for (int i = 0; i < 13; ++i) row[k++] = preserved[i];
            row[3] = (double)t;
            for (int i = 0; i < 4; ++i) row[k++] = tgt_pub[t][i];
            row[k++] = cfg->target_step_size / (double)e->tgt_cap[t];
            row[k++] = (double)e->tgt_cap[t];
            for (int w = 0; w < NW; ++w) row[k++] = (e->tgt_goal[t] == w) ? (double)e->tgt_weight[t] : 0.0;
            for (int w = 0; w < NW; ++w) row[k++] = (e->tgt_empty[t] >> w) & 1 ? 1.0 : 0.0;
            for (int c = 0; c < nc; ++c) for (int i = 0; i < 7; ++i) row[k++] = e->m_tc[t][c] ? cam_pub[c][i] : 0.0;
            for (int o = 0; o < no; ++o) {
                int m = e->m_to[t][o];
                row[k++] = m ? e->obs_x[o] : 0.0; row[k++] = m ? e->obs_y[o] : 0.0;
                row[k++] = m ? e->obs_r[o] : 0.0; row[k++] = m ? 1.0 : 0.0;
            }
            for (int u = 0; u < nt; ++u) for (int i = 0; i < 5; ++i) row[k++] = e->m_tt[t][u] ? tgt_pub[u][i] : 0.0;
        }
    }
    /* coverage statistics (environment.py:966-979) */
    int tracked = 0, with_bounty = 0, both = 0;
    for (int t = 0; t < nt; ++t) {
        tracked += e->tracked[t];
        if (e->tgt_bounty[t] > 0) { with_bounty++; both += e->tracked[t]; }
    }
    e->coverage = (double)tracked / (double)nt;
    e->real_coverage = with_bounty > 0 ? (double)both / (double)with_bounty : 0.0;
    e->transport = e->delivered > 0 ? e->delayed_ep_reward / (S->reward_scale * e->delivered) : 0.0;
}

static void refresh_derived(const Oracle* S, Env* e) {
    const MateConfig* cfg = &S->cfg;
    double area_product = cfg->camera_min_viewing_angle * (cfg->camera_max_sight_range * cfg->camera_max_sight_range);
    for (int c = 0; c < cfg->num_cameras; ++c) {
        e->cam_rs[c] = sqrt(area_product / e->cam_theta[c]);
        build_fov(S, e, c);
    }
}

/* ------------------------------------------------------------------ reset (A11) */

static int overlap_plain(double x0, double y0, double r0, double x1, double y1, double r1, double min_distance) {
    /* Entity.overlap (mate/entities.py:96-100) */
    return norm2(x0 - x1, y0 - y1) * (1.0 + 1e-6) < r0 + r1 + min_distance;
}

/* MultiAgentTracking.reset (mate/environment.py:679-834) on Philox streams. */
static void env_reset(const Oracle* S, Env* e, uint64_t seed, int env_index) {
    const MateConfig* cfg = &S->cfg;
    const int nc = cfg->num_cameras, nt = cfg->num_targets, no = cfg->num_obstacles;
    RngKey key = {seed, (uint32_t)(S->env_index_base + env_index), (uint32_t)e->episode_id};
    int perm_c[MAXC], perm_t[MAXT], perm_o[MAXO];
    for (int i = 0; i < nc; ++i) perm_c[i] = i;
    for (int i = 0; i < nt; ++i) perm_t[i] = i;
    for (int i = 0; i < no; ++i) perm_o[i] = i;
    if (cfg->shuffle_entities) { /* environment.py:707-710; RandomState.shuffle = Fisher-Yates from the top */
        for (int i = nc - 1; i >= 1; --i) { int j = (int)rng_below(&key, STREAM_SHUFFLE_CAM, (uint32_t)i, (uint32_t)(i + 1)); int s = perm_c[i]; perm_c[i] = perm_c[j]; perm_c[j] = s; }
        for (int i = nt - 1; i >= 1; --i) { int j = (int)rng_below(&key, STREAM_SHUFFLE_TGT, (uint32_t)i, (uint32_t)(i + 1)); int s = perm_t[i]; perm_t[i] = perm_t[j]; perm_t[j] = s; }
        for (int i = no - 1; i >= 1; --i) { int j = (int)rng_below(&key, STREAM_SHUFFLE_OBS, (uint32_t)i, (uint32_t)(i + 1)); int s = perm_o[i]; perm_o[i] = perm_o[j]; perm_o[j] = s; }
    }
    /* capacities (environment.py:712-722) */
    for (int t = 0; t < nt; ++t) e->tgt_cap[t] = 1;
    if (cfg->num_high_capacity_targets > 0) {
        if (cfg->shuffle_entities) { /* choice(Nt, size=k, replace=False): partial Fisher-Yates */
            int idx[MAXT];
            for (int i = 0; i < nt; ++i) idx[i] = i;
            for (int i = 0; i < cfg->num_high_capacity_targets; ++i) {
                int j = i + (int)rng_below(&key, STREAM_CAPACITY, (uint32_t)i, (uint32_t)(nt - i));
                int s = idx[i]; idx[i] = idx[j]; idx[j] = s;
                e->tgt_cap[idx[i]] = 2;
            }
        } else {
            for (int i = 0; i < cfg->num_high_capacity_targets; ++i) e->tgt_cap[i] = 2;
        }
    }
    /* rejection placement (environment.py:724-737): cameras, obstacles, targets */
    double px[NW + MAXC + MAXO + MAXT], py[NW + MAXC + MAXO + MAXT], pr[NW + MAXC + MAXO + MAXT], prs[NW + MAXC + MAXO + MAXT];
    int pcam[NW + MAXC + MAXO + MAXT];
    int placed = 0;
    for (int w = 0; w < NW; ++w) { px[placed] = WAREHOUSES[w][0]; py[placed] = WAREHOUSES[w][1]; pr[placed] = 0.75 * WAREHOUSE_RADIUS; prs[placed] = 0; pcam[placed] = 0; ++placed; }
    const double area_product = cfg->camera_min_viewing_angle * (cfg->camera_max_sight_range * cfg->camera_max_sight_range);
    int serial = 0;
    for (int kind = 0; kind < 3; ++kind) {
        int count = kind == 0 ? nc : (kind == 1 ? no : nt);
        for (int i = 0; i < count; ++i, ++serial) {
            const double* range = kind == 0 ? S->cam_ranges[perm_c[i]] : (kind == 1 ? S->obs_ranges[perm_o[i]] : S->tgt_ranges[perm_t[i]]);
            double min_distance = kind == 2 ? 0.0 : cfg->target_step_size;
            double x = 0, y = 0, radius = kind == 0 ? cfg->camera_radius : 0.0, phi = 0, theta = 0, rs = 0;
            int ok = 0;
            for (int attempt = 0; attempt < NUM_RESET_RETRIES && !ok; ++attempt) {
                uint32_t base = ((uint32_t)serial * NUM_RESET_RETRIES + (uint32_t)attempt) * 8u;
                if (kind == 1) /* Obstacle.reset: radius first (entities.py:150-152) */
                    radius = cfg->obstacle_radius_low + (cfg->obstacle_radius_high - cfg->obstacle_radius_low) * rng_u01(&key, STREAM_PLACE, base + 2);
                /* Entity.reset (entities.py:60-65) */
                x = range[0] + (range[1] - range[0]) * rng_u01(&key, STREAM_PLACE, base + 0);
                y = range[2] + (range[3] - range[2]) * rng_u01(&key, STREAM_PLACE, base + 1);
                double lim = TERRAIN_SIZE - 1.2 * radius;
                x = fmin(fmax(x, -lim), lim);
                y = fmin(fmax(y, -lim), lim);
                if (kind == 0) { /* Camera.reset (entities.py:326-334) */
                    uint32_t nrot = (uint32_t)(360.0 / cfg->camera_rotation_step);
                    phi = normalize_angle(cfg->camera_rotation_step * (double)rng_below(&key, STREAM_PLACE, base + 3, nrot));
                    theta = cfg->camera_min_viewing_angle + (180.0 - cfg->camera_min_viewing_angle) * rng_u01(&key, STREAM_PLACE, base + 4);
                    rs = sqrt(area_product / theta);
                }
                ok = 1;
                for (int q = 0; q < placed && ok; ++q) {
                    if (overlap_plain(x, y, radius, px[q], py[q], pr[q], min_distance)) ok = 0;
                    else if (kind == 0 && pcam[q]) { /* Camera.overlap (entities.py:484-489) */
                        if (norm2(x - px[q], y - py[q]) < 0.1 * fmin(rs, prs[q])) ok = 0;
                    }
                }
            }
            if (!ok && kind == 1) radius = 0.0; /* environment.py:734-736 */
            px[placed] = x; py[placed] = y; pr[placed] = radius; prs[placed] = rs; pcam[placed] = kind == 0; ++placed;
            if (kind == 0) { e->cam_x[i] = x; e->cam_y[i] = y; e->cam_phi[i] = phi; e->cam_theta[i] = theta; e->cam_rs[i] = rs; }
            else if (kind == 1) { e->obs_x[i] = x; e->obs_y[i] = y; e->obs_r[i] = radius; }
            else { e->tgt_x[i] = x; e->tgt_y[i] = y; }
        }
    }
    for (int t = 0; t < nt; ++t) { e->tgt_goal[t] = -1; e->tgt_weight[t] = 0; e->tgt_bounty[t] = 0; e->tgt_empty[t] = 0; e->tgt_colliding[t] = 0; e->tgt_done[t] = 0; e->tgt_orient[t] = 0.0; }
    /* cargo table (environment.py:768-775) */
    memset(e->remaining, 0, sizeof(e->remaining));
    uint32_t draw = 0;
    for (;;) {
        int all_rows = 1;
        for (int w = 0; w < NW; ++w) { int any = 0; for (int g = 0; g < NW; ++g) any |= e->remaining[w][g] > 0; all_rows &= any; }
        if (all_rows) break;
        for (int i = 0; i < cfg->num_cargoes_per_target * nt; ++i, ++draw) {
            uint32_t w[4];
            rng_words(&key, STREAM_CARGO, draw, w);
            int sender = (int)(((uint64_t)w[0] * NW) >> 32);
            int recipient = (int)(((uint64_t)w[1] * (NW - 1)) >> 32);
            if (recipient >= sender) recipient += 1; /* choice(4, size=2, replace=False) */
            e->remaining[sender][recipient] += 1;
        }
    }
    for (int g = 0; g < NW; ++g) { e->awaiting[g] = 0; for (int w = 0; w < NW; ++w) e->awaiting[g] += e->remaining[w][g]; }
    e->delivered = 0; e->ep_reward = 0.0; e->delayed_ep_reward = 0.0; e->episode_step = 0; e->coverage_sum = 0.0;
    refresh_derived(S, e);
}

/* second half of reset: _update_view, initial _assign_goals, start-with-cargo assignment
 * (environment.py:766, 784-812) */
static void env_reset_finish(const Oracle* S, Env* e, uint64_t seed, int env_index) {
    const MateConfig* cfg = &S->cfg;
    const int nt = cfg->num_targets;
    RngKey key = {seed, (uint32_t)(S->env_index_base + env_index), (uint32_t)e->episode_id};
    update_view(S, e, NULL, &key, 0);
    double r, d;
    assign_goals(S, e, NULL, &key, 0, &r, &d);
    for (int t = 0; t < nt; ++t) e->tgt_done[t] = 0;
    e->delivered = 0; e->ep_reward = 0.0; e->delayed_ep_reward = 0.0;
    if (cfg->targets_start_with_cargoes) {
        for (int t = 0; t < nt; ++t) {
            if (e->tgt_goal[t] >= 0) continue;
            int perm[NW] = {0, 1, 2, 3}; /* np_random.permutation(4) */
            for (int i = NW - 1; i >= 1; --i) {
                int j = (int)rng_below(&key, STREAM_INIT_GOAL, (uint32_t)(t * 8 + i), (uint32_t)(i + 1));
                int s = perm[i]; perm[i] = perm[j]; perm[j] = s;
            }
            for (int k = 0; k < NW; ++k) {
                int w = perm[k];
                int cand[NW], ncand = 0;
                for (int g = 0; g < NW; ++g) if (e->remaining[w][g] > 0) cand[ncand++] = g;
                if (ncand == 0) continue;
                int goal = cand[rng_below(&key, STREAM_INIT_GOAL, (uint32_t)(t * 8 + 4), (uint32_t)ncand)];
                int rem = e->remaining[w][goal];
                int weight = e->tgt_cap[t] < rem ? e->tgt_cap[t] : rem;
                e->remaining[w][goal] -= weight;
                e->tgt_weight[t] = weight;
                e->tgt_bounty[t] = (int)(weight * S->bounty_scale);
                e->tgt_goal[t] = goal;
                break;
            }
        }
    }
}


/* ------------------------------------------------------------------ threading (pthreads) */
#include <pthread.h>

typedef void (*RangeFn)(void* ctx, int begin, int end, int worker);
typedef struct { RangeFn fn; void* ctx; int begin, end, worker; } RangeJob;
static void* range_trampoline(void* p) { RangeJob* j = (RangeJob*)p; j->fn(j->ctx, j->begin, j->end, j->worker); return NULL; }

static void parallel_ranges(int n, int num_threads, RangeFn fn, void* ctx) {
    if (num_threads > n) num_threads = n;
    if (num_threads <= 1) { fn(ctx, 0, n, 0); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)num_threads);
    RangeJob* jobs = (RangeJob*)malloc(sizeof(RangeJob) * (size_t)num_threads);
    for (int i = 0; i < num_threads; ++i) {
        jobs[i].fn = fn; jobs[i].ctx = ctx; jobs[i].worker = i;
        jobs[i].begin = (int)((int64_t)n * i / num_threads);
        jobs[i].end = (int)((int64_t)n * (i + 1) / num_threads);
        pthread_create(&th[i], NULL, range_trampoline, &jobs[i]);
    }
    for (int i = 0; i < num_threads; ++i) pthread_join(th[i], NULL);
    free(th); free(jobs);
}

/* ------------------------------------------------------------------ exported API */

static void write_aux(const Oracle* S, const Env* e, int b, const MateStepAux* aux) {
    if (!aux) return;
    const MateConfig* cfg = &S->cfg;
    const int nc = cfg->num_cameras, nt = cfg->num_targets, no = cfg->num_obstacles;
    for (int c = 0; c < nc; ++c) {
        for (int t = 0; t < nt; ++t) if (aux->mask_ct) aux->mask_ct[((size_t)b * nc + c) * nt + t] = e->m_ct[c][t];
        for (int d = 0; d < nc; ++d) if (aux->mask_cc) aux->mask_cc[((size_t)b * nc + c) * nc + d] = e->m_cc[c][d];
        for (int o = 0; o < no; ++o) if (aux->mask_co) aux->mask_co[((size_t)b * nc + c) * no + o] = e->m_co[c][o];
    }
    for (int t = 0; t < nt; ++t) {
        for (int c = 0; c < nc; ++c) if (aux->mask_tc) aux->mask_tc[((size_t)b * nt + t) * nc + c] = e->m_tc[t][c];
        for (int o = 0; o < no; ++o) if (aux->mask_to) aux->mask_to[((size_t)b * nt + t) * no + o] = e->m_to[t][o];
        for (int u = 0; u < nt; ++u) if (aux->mask_tt) aux->mask_tt[((size_t)b * nt + t) * nt + u] = e->m_tt[t][u];
        if (aux->target_dones) aux->target_dones[(size_t)b * nt + t] = (uint8_t)e->tgt_done[t];
        if (aux->is_colliding) aux->is_colliding[(size_t)b * nt + t] = (uint8_t)e->tgt_colliding[t];
        if (aux->warehouse_dist) for (int w = 0; w < NW; ++w) aux->warehouse_dist[((size_t)b * nt + t) * NW + w] = (float)e->wh_dist[t][w];
    }
    if (aux->coverage) { aux->coverage[(size_t)b * 3 + 0] = (float)e->coverage; aux->coverage[(size_t)b * 3 + 1] = (float)e->real_coverage; aux->coverage[(size_t)b * 3 + 2] = (float)e->transport; }
    if (aux->num_delivered) aux->num_delivered[b] = e->delivered;
    if (aux->episode_step) aux->episode_step[b] = e->episode_step;
}

void* oracle_create(const MateConfig* cfg, int32_t num_envs, int64_t env_index_base) {
    if (!cfg || num_envs <= 0 || cfg->num_cameras > MAXC || cfg->num_targets > MAXT || cfg->num_targets < 1 || cfg->num_obstacles > MAXO) return NULL;
    Oracle* S = (Oracle*)calloc(1, sizeof(Oracle));
    S->cfg = *cfg;
    for (int i = 0; i < cfg->num_cameras; ++i) memcpy(S->cam_ranges[i], cfg->camera_location_ranges + 4 * i, 4 * sizeof(double));
    for (int i = 0; i < cfg->num_targets; ++i) memcpy(S->tgt_ranges[i], cfg->target_location_ranges + 4 * i, 4 * sizeof(double));
    for (int i = 0; i < cfg->num_obstacles; ++i) memcpy(S->obs_ranges[i], cfg->obstacle_location_ranges + 4 * i, 4 * sizeof(double));
    S->cfg.camera_location_ranges = S->cfg.target_location_ranges = S->cfg.obstacle_location_ranges = NULL;
    S->num_envs = num_envs;
    S->num_threads = 1;
    S->env_index_base = env_index_base;
    const int nc = cfg->num_cameras, nt = cfg->num_targets, no = cfg->num_obstacles;
    S->dc = 13 + 9 + 5 * nt + 4 * no + 7 * nc; /* constants.py:267-282 */
    S->dt = 13 + 14 + 7 * nc + 4 * no + 5 * nt; /* constants.py:285-300 */
    S->freight_scale = ceil(2.0 * TERRAIN_SIZE / cfg->target_step_size); /* environment.py:521-523 */
    S->bounty_scale = ceil(S->freight_scale * fmax(0.0, cfg->bounty_factor));
    S->reward_scale = S->freight_scale + S->bounty_scale;
    S->fov_cap = 360 + no * MAX_RAYS_PER_OBSTACLE + 2;
    S->envs = (Env*)calloc((size_t)num_envs, sizeof(Env));
    for (int b = 0; b < num_envs; ++b) {
        Env* e = &S->envs[b];
        for (int c = 0; c < nc; ++c) {
            e->fov_phi[c] = (double*)malloc(sizeof(double) * (size_t)S->fov_cap);
            e->fov_rho[c] = (double*)malloc(sizeof(double) * (size_t)S->fov_cap);
            e->cam_theta[c] = cfg->camera_min_viewing_angle;
        }
        for (int t = 0; t < nt; ++t) { e->tgt_cap[t] = 1; e->tgt_goal[t] = -1; }
    }
    return S;
}

void oracle_destroy(void* handle) {
    Oracle* S = (Oracle*)handle;
    if (!S) return;
    for (int b = 0; b < S->num_envs; ++b)
        for (int c = 0; c < S->cfg.num_cameras; ++c) { free(S->envs[b].fov_phi[c]); free(S->envs[b].fov_rho[c]); }
    free(S->envs);
    free(S);
}

typedef struct { Oracle* S; const MateStateView* v; } SetStateCtx;
static void set_state_range(void* p, int begin, int end, int worker) {
    (void)worker;
    Oracle* S = ((SetStateCtx*)p)->S;
    const MateStateView* v = ((SetStateCtx*)p)->v;
    const MateConfig* cfg = &S->cfg;
    const int nc = cfg->num_cameras, nt = cfg->num_targets, no = cfg->num_obstacles;
    for (int b = begin; b < end; ++b) {
        Env* e = &S->envs[b];
        for (int c = 0; c < nc; ++c) {
            if (v->cam_xy) { e->cam_x[c] = v->cam_xy[((size_t)b * nc + c) * 2]; e->cam_y[c] = v->cam_xy[((size_t)b * nc + c) * 2 + 1]; }
            if (v->cam_phi) e->cam_phi[c] = v->cam_phi[(size_t)b * nc + c];
            if (v->cam_theta) e->cam_theta[c] = v->cam_theta[(size_t)b * nc + c];
        }
        for (int t = 0; t < nt; ++t) {
            if (v->tgt_xy) { e->tgt_x[t] = v->tgt_xy[((size_t)b * nt + t) * 2]; e->tgt_y[t] = v->tgt_xy[((size_t)b * nt + t) * 2 + 1]; }
            if (v->tgt_capacity) e->tgt_cap[t] = v->tgt_capacity[(size_t)b * nt + t];
            if (v->tgt_goal) e->tgt_goal[t] = v->tgt_goal[(size_t)b * nt + t];
            if (v->tgt_weight) e->tgt_weight[t] = v->tgt_weight[(size_t)b * nt + t];
            if (v->tgt_bounty) e->tgt_bounty[t] = v->tgt_bounty[(size_t)b * nt + t];
            if (v->tgt_empty_bits) e->tgt_empty[t] = v->tgt_empty_bits[(size_t)b * nt + t];
        }
        for (int o = 0; o < no; ++o)
            if (v->obs_xyr) { e->obs_x[o] = v->obs_xyr[((size_t)b * no + o) * 3]; e->obs_y[o] = v->obs_xyr[((size_t)b * no + o) * 3 + 1]; e->obs_r[o] = v->obs_xyr[((size_t)b * no + o) * 3 + 2]; }
        if (v->remaining) for (int i = 0; i < NW * NW; ++i) e->remaining[i / NW][i % NW] = v->remaining[(size_t)b * NW * NW + i];
        if (v->awaiting) for (int i = 0; i < NW; ++i) e->awaiting[i] = v->awaiting[(size_t)b * NW + i];
        if (v->num_delivered) e->delivered = v->num_delivered[b];
        if (v->episode_step) e->episode_step = v->episode_step[b];
        if (v->episode_id) e->episode_id = v->episode_id[b];
        if (v->episode_reward) { e->ep_reward = v->episode_reward[(size_t)b * 2]; e->delayed_ep_reward = v->episode_reward[(size_t)b * 2 + 1]; }
        refresh_derived(S, e);
    }
}
int oracle_set_state(void* handle, const MateStateView* v) {
    SetStateCtx ctx = {(Oracle*)handle, v};
    parallel_ranges(ctx.S->num_envs, ctx.S->num_threads, set_state_range, &ctx);
    return 0;
}

int oracle_get_state(void* handle, MateStateView* v) {
    Oracle* S = (Oracle*)handle;
    const MateConfig* cfg = &S->cfg;
    const int nc = cfg->num_cameras, nt = cfg->num_targets, no = cfg->num_obstacles;
    for (int b = 0; b < S->num_envs; ++b) {
        const Env* e = &S->envs[b];
        for (int c = 0; c < nc; ++c) {
            if (v->cam_xy) { v->cam_xy[((size_t)b * nc + c) * 2] = e->cam_x[c]; v->cam_xy[((size_t)b * nc + c) * 2 + 1] = e->cam_y[c]; }
            if (v->cam_phi) v->cam_phi[(size_t)b * nc + c] = e->cam_phi[c];
            if (v->cam_theta) v->cam_theta[(size_t)b * nc + c] = e->cam_theta[c];
        }
        for (int t = 0; t < nt; ++t) {
            if (v->tgt_xy) { v->tgt_xy[((size_t)b * nt + t) * 2] = e->tgt_x[t]; v->tgt_xy[((size_t)b * nt + t) * 2 + 1] = e->tgt_y[t]; }
            if (v->tgt_capacity) v->tgt_capacity[(size_t)b * nt + t] = e->tgt_cap[t];
            if (v->tgt_goal) v->tgt_goal[(size_t)b * nt + t] = e->tgt_goal[t];
            if (v->tgt_weight) v->tgt_weight[(size_t)b * nt + t] = e->tgt_weight[t];
            if (v->tgt_bounty) v->tgt_bounty[(size_t)b * nt + t] = e->tgt_bounty[t];
            if (v->tgt_empty_bits) v->tgt_empty_bits[(size_t)b * nt + t] = e->tgt_empty[t];
        }
        for (int o = 0; o < no; ++o)
            if (v->obs_xyr) { v->obs_xyr[((size_t)b * no + o) * 3] = e->obs_x[o]; v->obs_xyr[((size_t)b * no + o) * 3 + 1] = e->obs_y[o]; v->obs_xyr[((size_t)b * no + o) * 3 + 2] = e->obs_r[o]; }
        if (v->remaining) for (int i = 0; i < NW * NW; ++i) v->remaining[(size_t)b * NW * NW + i] = e->remaining[i / NW][i % NW];
        if (v->awaiting) for (int i = 0; i < NW; ++i) v->awaiting[(size_t)b * NW + i] = e->awaiting[i];
        if (v->num_delivered) v->num_delivered[b] = e->delivered;
        if (v->episode_step) v->episode_step[b] = e->episode_step;
        if (v->episode_id) v->episode_id[b] = e->episode_id;
        if (v->episode_reward) { v->episode_reward[(size_t)b * 2] = e->ep_reward; v->episode_reward[(size_t)b * 2 + 1] = e->delayed_ep_reward; }
    }
    return 0;
}

/* number of (phi, rho) samples of camera c's FOV polyline in env b; copies up to cap */
int oracle_get_fov(void* handle, int b, int c, double* phi, double* rho, int cap) {
    Oracle* S = (Oracle*)handle;
    const Env* e = &S->envs[b];
    int n = e->fov_n[c];
    for (int i = 0; i < n && i < cap; ++i) { phi[i] = e->fov_phi[c][i]; rho[i] = e->fov_rho[c][i]; }
    return n;
}

/* joint_observation of the current state (after set_state / reset). transmit: host [B,Nc,Nt] or NULL */
typedef struct {
    Oracle* S; const double* cam_act; const double* tgt_act; const uint8_t* transmit; const int8_t* goal_choice;
    const uint8_t* env_mask; uint64_t seed; uint32_t flags; double* cam_obs; double* tgt_obs; double* rewards;
    uint8_t* done; const MateStepAux* aux; double (*stats)[16];
} CallCtx;

static void observe_range(void* p, int begin, int end, int worker) {
    (void)worker;
    CallCtx* C = (CallCtx*)p;
    Oracle* S = C->S;
    const uint8_t* transmit = C->transmit; uint64_t seed = C->seed;
    double* cam_obs = C->cam_obs; double* tgt_obs = C->tgt_obs; const MateStepAux* aux = C->aux;
    const MateConfig* cfg = &S->cfg;
    const int nc = cfg->num_cameras, nt = cfg->num_targets;
    for (int b = begin; b < end; ++b) {
        Env* e = &S->envs[b];
        RngKey key = {seed, (uint32_t)(S->env_index_base + b), (uint32_t)e->episode_id};
        update_view(S, e, transmit ? transmit + (size_t)b * nc * nt : NULL, &key, e->episode_step);
        joint_observation(S, e, cam_obs ? cam_obs + (size_t)b * nc * S->dc : NULL, tgt_obs ? tgt_obs + (size_t)b * nt * S->dt : NULL);
        write_aux(S, e, b, aux);
    }
}
int oracle_observe(void* handle, const uint8_t* transmit, uint64_t seed, double* cam_obs, double* tgt_obs,
                   const MateStepAux* aux) {
    CallCtx C; memset(&C, 0, sizeof(C));
    C.S = (Oracle*)handle; C.transmit = transmit; C.seed = seed; C.cam_obs = cam_obs; C.tgt_obs = tgt_obs; C.aux = aux;
    parallel_ranges(C.S->num_envs, C.S->num_threads, observe_range, &C);
    return 0;
}

static void reset_range(void* p, int begin, int end, int worker) {
    (void)worker;
    CallCtx* C = (CallCtx*)p;
    Oracle* S = C->S;
    const uint8_t* env_mask = C->env_mask; uint64_t seed = C->seed;
    double* cam_obs = C->cam_obs; double* tgt_obs = C->tgt_obs;
    const MateConfig* cfg = &S->cfg;
    const int nc = cfg->num_cameras, nt = cfg->num_targets;
    for (int b = begin; b < end; ++b) {
        Env* e = &S->envs[b];
        if (!env_mask || env_mask[b]) {
            e->episode_id += 1;
            env_reset(S, e, seed, b);
            env_reset_finish(S, e, seed, b);
        } else {
            RngKey key = {seed, (uint32_t)(S->env_index_base + b), (uint32_t)e->episode_id};
            update_view(S, e, NULL, &key, e->episode_step);
        }
        joint_observation(S, e, cam_obs ? cam_obs + (size_t)b * nc * S->dc : NULL, tgt_obs ? tgt_obs + (size_t)b * nt * S->dt : NULL);
    }
}
int oracle_reset(void* handle, const uint8_t* env_mask, uint64_t seed, double* cam_obs, double* tgt_obs) {
    CallCtx C; memset(&C, 0, sizeof(C));
    C.S = (Oracle*)handle; C.env_mask = env_mask; C.seed = seed; C.cam_obs = cam_obs; C.tgt_obs = tgt_obs;
    parallel_ranges(C.S->num_envs, C.S->num_threads, reset_range, &C);
    return 0;
}

/* MultiAgentTracking.step (mate/environment.py:590-676) for every env.
 * Actions are float64 [B,Nc,2] / [B,Nt,2]; observations float64 (may be NULL to skip
 * writing them, the packing work is still done into a scratch row). */
static void step_range(void* p, int begin, int end, int worker) {
    CallCtx* C = (CallCtx*)p;
    Oracle* S = C->S;
    const MateConfig* cfg = &S->cfg;
    const int nc = cfg->num_cameras, nt = cfg->num_targets;
    const double* cam_act = C->cam_act; const double* tgt_act = C->tgt_act;
    const uint8_t* transmit = C->transmit; const int8_t* goal_choice = C->goal_choice;
    const uint64_t seed = C->seed; const uint32_t flags = C->flags;
    double* scratch_c = (double*)malloc(sizeof(double) * (size_t)(nc * S->dc + 1));
    double* scratch_t = (double*)malloc(sizeof(double) * (size_t)(nt * S->dt + 1));
    double* st = C->stats[worker];
    for (int b = begin; b < end; ++b) {
        Env* e = &S->envs[b];
        RngKey key = {seed, (uint32_t)(S->env_index_base + b), (uint32_t)e->episode_id};
        /* _simulate (environment.py:1326-1354) */
        for (int c = 0; c < nc; ++c)
            camera_simulate(cfg, e, c, cam_act[((size_t)b * nc + c) * 2], cam_act[((size_t)b * nc + c) * 2 + 1]);
        for (int t = 0; t < nt; ++t)
            target_simulate(cfg, e, t, tgt_act[((size_t)b * nt + t) * 2], tgt_act[((size_t)b * nt + t) * 2 + 1]);
        update_view(S, e, transmit ? transmit + (size_t)b * nc * nt : NULL, &key, e->episode_step + 1);
        double r, delayed;
        assign_goals(S, e, goal_choice ? goal_choice + (size_t)b * nt : NULL, &key, e->episode_step + 1, &r, &delayed);
        e->ep_reward += r;
        e->delayed_ep_reward += delayed;
        double* co = C->cam_obs ? C->cam_obs + (size_t)b * nc * S->dc : scratch_c;
        double* to = C->tgt_obs ? C->tgt_obs + (size_t)b * nt * S->dt : scratch_t;
        joint_observation(S, e, co, to);
        if (cfg->reward_sparse) r = delayed; /* environment.py:618-619 */
        e->episode_step += 1;
        e->coverage_sum += e->coverage;
        int any_awaiting = 0;
        for (int w = 0; w < NW; ++w) any_awaiting |= e->awaiting[w] != 0;
        int done = !(e->episode_step <= cfg->max_episode_steps && any_awaiting); /* environment.py:630-632 */
        if (C->rewards) { C->rewards[(size_t)b * 2] = -r; C->rewards[(size_t)b * 2 + 1] = r; }
        if (C->done) C->done[b] = (uint8_t)done;
        write_aux(S, e, b, C->aux);
        st[5] += 1.0;
        if (done) {
            st[0] += 1.0; st[1] += e->ep_reward; st[2] += e->episode_step; st[3] += e->delivered;
            st[4] += e->coverage_sum / (double)e->episode_step;
            if (flags & MATE_STEP_AUTO_RESET) {
                e->episode_id += 1;
                env_reset(S, e, seed, b);
                env_reset_finish(S, e, seed, b);
                joint_observation(S, e, co, to);
            }
        }
    }
    free(scratch_c);
    free(scratch_t);
}

int oracle_step(void* handle, const double* cam_act, const double* tgt_act, const uint8_t* transmit,
                const int8_t* goal_choice, uint64_t seed, uint32_t flags, double* cam_obs,
                double* tgt_obs, double* rewards, uint8_t* done_out, const MateStepAux* aux) {
    Oracle* S = (Oracle*)handle;
    int nthreads = S->num_threads < 1 ? 1 : S->num_threads;
    double (*stats)[16] = (double (*)[16])calloc((size_t)nthreads, sizeof(double[16]));
    CallCtx C; memset(&C, 0, sizeof(C));
    C.S = S; C.cam_act = cam_act; C.tgt_act = tgt_act; C.transmit = transmit; C.goal_choice = goal_choice;
    C.seed = seed; C.flags = flags; C.cam_obs = cam_obs; C.tgt_obs = tgt_obs; C.rewards = rewards;
    C.done = done_out; C.aux = aux; C.stats = stats;
    parallel_ranges(S->num_envs, nthreads, step_range, &C);
    for (int w = 0; w < nthreads; ++w) for (int i = 0; i < 16; ++i) S->stats[i] += stats[w][i];
    free(stats);
    return 0;
}

void oracle_set_threads(void* handle, int num_threads) { ((Oracle*)handle)->num_threads = num_threads < 1 ? 1 : num_threads; }

int oracle_episode_stats(void* handle, double* out16, int reset_after) {
    Oracle* S = (Oracle*)handle;
    for (int i = 0; i < 16; ++i) out16[i] = S->stats[i];
    if (reset_after) memset(S->stats, 0, sizeof(S->stats));
    return 0;
}

void oracle_obs_dims(void* handle, int32_t* dc, int32_t* dt) {
    Oracle* S = (Oracle*)handle;
    *dc = S->dc; *dt = S->dt;
}
