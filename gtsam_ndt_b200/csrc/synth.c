/*
 * Synthetic 2D laser world for tests and benchmarks (host only, no CUDA, no oracle).
 *
 * SURVEY.md section 8(d) "Synthetic inputs": an axis-aligned room with box obstacles inside the
 * 200 x 200 m extent, scans by exact ray casting, Gaussian range noise, SplitMix64 RNG.
 * There is no reference data set (the reference mount is /root/reference/README.md:1 only),
 * so this generator IS the data definition. Everything is a pure function of
 * (seed, scan index, beam index): output does not depend on thread count or call order.
 *
 * World: outer walls at +-95 m; `nboxes` candidate boxes (centre U[-90,90]^2, half sizes U[1,4] m)
 * kept only if they stay 3 m clear of the trajectory, a circle of radius 60 m.
 * Pose i of m: a = 2*pi*i/m, position 60*(cos a, sin a), heading a + pi/2 + 0.3*sin(5a).
 * Output is LaserScan-shaped: ranges[scan][beam] f32, 0 = no return (SPEC.md section 8).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SYNTH_MAX_BOXES 32768
#define WALL 95.0
#define TRAJ_R 60.0

typedef struct { double x0, y0, x1, y1; } box_t;

#define GRID_N 20          /* broad phase: 20 x 20 buckets of 10 m over [-100, 100]^2 */
#define GRID_CELL 10.0
#define GRID_MAX 512

typedef struct {
    int nboxes;
    box_t box[SYNTH_MAX_BOXES];
    int bucket_n[GRID_N][GRID_N];
    int bucket[GRID_N][GRID_N][GRID_MAX];
} world_t;

/* Shape of the candidate boxes: half sizes U[h_min, h_min + h_span] m, kept `clear` m away from the trajectory. The
 * defaults are the SURVEY 8(d) room (about 230 boxes survive of 400 candidates: 14 k occupied 0.25 m cells); the "dense"
 * world of bench.py (--world dense) uses many small boxes so that a third of the 640 k map cells are occupied. */
static double g_hmin = 1.0, g_hspan = 3.0, g_clear = 3.0;
/* threads of the ray-casting loops (torchrun exports OMP_NUM_THREADS=1 to every rank: the benchmark sets its share of the
 * host cores explicitly); n <= 0 leaves the OpenMP default */
void synth_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void synth_set_world(double h_min, double h_span, double clear)
{
    g_hmin = h_min; g_hspan = h_span; g_clear = clear;
}

static uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double u01(uint64_t *s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

static void build_world(world_t *w, uint64_t seed, int ncand)
{
    uint64_t s = seed;
    w->nboxes = 0;
    if (ncand > SYNTH_MAX_BOXES) ncand = SYNTH_MAX_BOXES;
    for (int i = 0; i < ncand; ++i) {
        double cx = -90.0 + 180.0 * u01(&s), cy = -90.0 + 180.0 * u01(&s);
        double hx = g_hmin + g_hspan * u01(&s), hy = g_hmin + g_hspan * u01(&s);
        double dist = fabs(sqrt(cx * cx + cy * cy) - TRAJ_R);
        if (dist < sqrt(hx * hx + hy * hy) + g_clear) continue;
        box_t b = {cx - hx, cy - hy, cx + hx, cy + hy};
        if (b.x0 < -WALL + 0.5 || b.x1 > WALL - 0.5 || b.y0 < -WALL + 0.5 || b.y1 > WALL - 0.5) continue;
        w->box[w->nboxes++] = b;
    }
    for (int gy = 0; gy < GRID_N; ++gy) for (int gx = 0; gx < GRID_N; ++gx) w->bucket_n[gy][gx] = 0;
    for (int i = 0; i < w->nboxes; ++i) {
        const box_t *b = &w->box[i];
        int x0 = (int)floor((b->x0 + 100.0) / GRID_CELL), x1 = (int)floor((b->x1 + 100.0) / GRID_CELL);
        int y0 = (int)floor((b->y0 + 100.0) / GRID_CELL), y1 = (int)floor((b->y1 + 100.0) / GRID_CELL);
        for (int gy = y0; gy <= y1; ++gy) for (int gx = x0; gx <= x1; ++gx)
            if (gx >= 0 && gx < GRID_N && gy >= 0 && gy < GRID_N && w->bucket_n[gy][gx] < GRID_MAX)
                w->bucket[gy][gx][w->bucket_n[gy][gx]++] = i;
    }
}

/* distance along (dx,dy) from (px,py) to the first surface */
static double cast(const world_t *w, double px, double py, double dx, double dy)
{
    double idx = 1.0 / dx, idy = 1.0 / dy; /* +-inf for axis-parallel rays is fine */
    /* outer walls: we are inside, take the exit distance */
    double tx = ((dx > 0 ? WALL : -WALL) - px) * idx;
    double ty = ((dy > 0 ? WALL : -WALL) - py) * idy;
    double best = fmin(dx != 0 ? tx : INFINITY, dy != 0 ? ty : INFINITY);
    /* walk the buckets along the ray (DDA); the first bucket whose nearest hit lies inside it ends the walk.
       The result is the minimum over all boxes, exactly as a brute-force loop would give. */
    int gx = (int)floor((px + 100.0) / GRID_CELL), gy = (int)floor((py + 100.0) / GRID_CELL);
    int sx = dx > 0 ? 1 : -1, sy = dy > 0 ? 1 : -1;
    double tmx = dx != 0 ? (((gx + (dx > 0)) * GRID_CELL - 100.0) - px) * idx : INFINITY;
    double tmy = dy != 0 ? (((gy + (dy > 0)) * GRID_CELL - 100.0) - py) * idy : INFINITY;
    double tdx = dx != 0 ? GRID_CELL * fabs(idx) : INFINITY, tdy = dy != 0 ? GRID_CELL * fabs(idy) : INFINITY;
    double tenter = 0.0;
    while (gx >= 0 && gx < GRID_N && gy >= 0 && gy < GRID_N && tenter < best) {
        double texit = fmin(tmx, tmy);
        for (int k = 0; k < w->bucket_n[gy][gx]; ++k) {
            const box_t *b = &w->box[w->bucket[gy][gx][k]];
            double t0x = (b->x0 - px) * idx, t1x = (b->x1 - px) * idx;
            double t0y = (b->y0 - py) * idy, t1y = (b->y1 - py) * idy;
            double tnx = fmin(t0x, t1x), tfx = fmax(t0x, t1x);
            double tny = fmin(t0y, t1y), tfy = fmax(t0y, t1y);
            double tn = fmax(tnx, tny), tf = fmin(tfx, tfy);
            if (tn <= tf && tn > 0.0 && tn < best) best = tn;
        }
        if (best <= texit) break; /* the nearest hit so far is inside the buckets already visited */
        tenter = texit;
        if (tmx < tmy) { gx += sx; tmx += tdx; } else { gy += sy; tmy += tdy; }
    }
    return best;
}

void synth_pose(int64_t i, int64_t m, double pose[3])
{
    const double TWO_PI = 6.283185307179586476925286766559;
    double a = TWO_PI * (double)i / (double)m;
    pose[0] = TRAJ_R * cos(a); pose[1] = TRAJ_R * sin(a);
    pose[2] = a + 0.5 * 3.14159265358979323846 + 0.3 * sin(5.0 * a);
}

/*
 * ranges[nscans*nbeams] f32 (0 = dropped), poses[nscans*3] f64 = true sensor poses.
 * Scan j is trajectory pose (first + j*step) of `traj_len`. Noise: N(0, sigma) on the range,
 * Box-Muller from a SplitMix64 stream keyed by (noise_seed, trajectory index, beam); `seed` fixes the world.
 * max_range <= 0: unlimited (the outer walls always return).
 */
int synth_scans(uint64_t seed, uint64_t noise_seed, int ncand_boxes, int64_t traj_len, int64_t first, int64_t step,
                int nscans, int nbeams, double angle_min, double angle_inc,
                double max_range, double sigma, float *ranges, double *poses)
{
    world_t *w = (world_t *)malloc(sizeof(world_t));
    if (!w) return 2;
    build_world(w, seed, ncand_boxes);
    const double TWO_PI = 6.283185307179586476925286766559;
    #pragma omp parallel for schedule(dynamic, 8)
    for (int j = 0; j < nscans; ++j) {
        int64_t ti = first + (int64_t)j * step;
        double pose[3];
        synth_pose(ti, traj_len, pose);
        poses[3 * j] = pose[0]; poses[3 * j + 1] = pose[1]; poses[3 * j + 2] = pose[2];
        for (int k = 0; k < nbeams; ++k) {
            double phi = pose[2] + (angle_min + (double)k * angle_inc);
            double r = cast(w, pose[0], pose[1], cos(phi), sin(phi));
            uint64_t s = noise_seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(ti + 1)) ^ (0x8CB92BA72F3D8DD7ull * (uint64_t)(k + 1));
            double u1 = u01(&s), u2 = u01(&s);
            if (u1 < 1e-300) u1 = 1e-300;
            r += sigma * sqrt(-2.0 * log(u1)) * cos(TWO_PI * u2);
            float rf = (float)r;
            if (!(rf > 0.0f) || (max_range > 0.0 && r > max_range)) rf = 0.0f;
            ranges[(size_t)j * nbeams + k] = rf;
        }
    }
    free(w);
    return 0;
}

/* uniform perturbations for initial guesses: out[n*3] in [-1,1), keyed by (seed, index) */
void synth_uniform3(uint64_t seed, int64_t first, int n, double *out)
{
    for (int i = 0; i < n; ++i) {
        uint64_t s = seed ^ (0xA0761D6478BD642Full * (uint64_t)(first + i + 1));
        for (int k = 0; k < 3; ++k) out[3 * i + k] = 2.0 * u01(&s) - 1.0;
    }
}

int synth_world_boxes(uint64_t seed, int ncand_boxes, double *boxes_out, int max_out)
{
    world_t *w = (world_t *)malloc(sizeof(world_t));
    if (!w) return -1;
    build_world(w, seed, ncand_boxes);
    int n = w->nboxes < max_out ? w->nboxes : max_out;
    for (int i = 0; i < n; ++i) {
        boxes_out[4 * i] = w->box[i].x0; boxes_out[4 * i + 1] = w->box[i].y0;
        boxes_out[4 * i + 2] = w->box[i].x1; boxes_out[4 * i + 3] = w->box[i].y1;
    }
    int total = w->nboxes;
    free(w);
    return total;
}
