// Deterministic synthetic LiDAR frame generator (test / benchmark tooling, not product code).
// Shapes follow SURVEY.md §8(d) configs 3-5: a spinning multi-beam sensor 1.73 m above a gently
// undulating ground, axis-aligned boxes and vertical cylinders as obstacles, nearest hit per ray,
// range noise, and rounding of every coordinate to 1 mm so that the reference's tie-heavy
// behaviour stays in play. RNG: SplitMix64 (scene parameters: sequential stream; per-ray noise:
// counter-based on the ray index), so the output is identical for any thread count.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace
{
struct SplitMix64
{
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return static_cast<double>(next() >> 11) * (1.0 / 9007199254740992.0); }
    double uniform(double a, double b) { return a + (b - a) * uniform(); }
};

inline uint64_t mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

inline double u01(uint64_t z) { return (static_cast<double>(z >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

struct Box
{
    double x0, x1, y0, y1, z0, z1;
};
struct Cyl
{
    double cx, cy, r, z0, z1;
};

inline double ground_z(double x, double y) { return -1.73 + 0.02 * std::sin(0.05 * x) * std::cos(0.05 * y); }

inline float quant_mm(double v) { return static_cast<float>(std::round(v * 1000.0) / 1000.0); }

struct Scene
{
    std::vector<Box> boxes;
    std::vector<Cyl> cyls;
};

Scene make_scene(uint64_t seed, uint32_t n_boxes, uint32_t n_cyls, double extent)
{
    SplitMix64 rng(seed * 0x100000001B3ull + 0x5851F42D4C957F2Dull);
    Scene sc;
    for (uint32_t i = 0; i < n_boxes; ++i)
    {
        const double w = rng.uniform(0.5, 5.0), l = rng.uniform(0.5, 5.0), h = rng.uniform(0.5, 3.0);
        const double cx = rng.uniform(-extent, extent), cy = rng.uniform(-extent, extent);
        if (std::fabs(cx) < 3.0 + w && std::fabs(cy) < 3.0 + l)
            continue; // keep the sensor outside every obstacle
        sc.boxes.push_back({cx - w / 2, cx + w / 2, cy - l / 2, cy + l / 2, -1.75, -1.73 + h});
    }
    for (uint32_t i = 0; i < n_cyls; ++i)
    {
        const double r = rng.uniform(0.1, 0.4), h = rng.uniform(1.0, 6.0);
        const double cx = rng.uniform(-extent, extent), cy = rng.uniform(-extent, extent);
        if (std::hypot(cx, cy) < 3.0)
            continue;
        sc.cyls.push_back({cx, cy, r, -1.75, -1.73 + h});
    }
    return sc;
}

// nearest positive hit distance of ray o + t*d (|d| = 1); returns max_range+1 when nothing is hit
double cast(const Scene &sc, const double o[3], const double d[3], double max_range)
{
    double best = max_range + 1.0;
    if (d[2] < -1e-6) // ground: fixed-point iteration on the undulating surface
    {
        double t = (-1.73 - o[2]) / d[2];
        for (int it = 0; it < 4; ++it)
            t = (ground_z(o[0] + t * d[0], o[1] + t * d[1]) - o[2]) / d[2];
        if (t > 0.5 && t < best)
            best = t;
    }
    for (const Box &b : sc.boxes)
    {
        double t0 = 0.0, t1 = best;
        const double lo[3] = {b.x0, b.y0, b.z0}, hi[3] = {b.x1, b.y1, b.z1};
        bool hit = true;
        for (int a = 0; a < 3 && hit; ++a)
        {
            if (std::fabs(d[a]) < 1e-12)
                hit = o[a] >= lo[a] && o[a] <= hi[a];
            else
            {
                double ta = (lo[a] - o[a]) / d[a], tb = (hi[a] - o[a]) / d[a];
                if (ta > tb)
                    std::swap(ta, tb);
                t0 = std::max(t0, ta);
                t1 = std::min(t1, tb);
                hit = t0 <= t1;
            }
        }
        if (hit && t0 > 0.5 && t0 < best)
            best = t0;
    }
    for (const Cyl &c : sc.cyls)
    {
        const double ox = o[0] - c.cx, oy = o[1] - c.cy;
        const double a = d[0] * d[0] + d[1] * d[1];
        if (a < 1e-12)
            continue;
        const double bq = ox * d[0] + oy * d[1];
        const double cq = ox * ox + oy * oy - c.r * c.r;
        const double disc = bq * bq - a * cq;
        if (disc < 0.0)
            continue;
        const double t = (-bq - std::sqrt(disc)) / a;
        if (t <= 0.5 || t >= best)
            continue;
        const double z = o[2] + t * d[2];
        if (z >= c.z0 && z <= c.z1)
            best = t;
    }
    return best;
}
} // namespace

extern "C"
{

// One spinning sensor at (sx, sy, 0). out: capacity beams*azimuth_steps records of 4 floats
// (x, y, z, intensity). Returns the number of returns written. Ray order: azimuth-major, beam-minor
// (like a spinning scanner's packet order).
uint32_t synth_sensor_frame(uint64_t scene_seed, uint64_t noise_seed, uint32_t beams, uint32_t azimuth_steps,
                            double elev_min_deg, double elev_max_deg, double sx, double sy, uint32_t n_boxes,
                            uint32_t n_cyls, float *out, uint32_t n_threads)
{
    const Scene sc = make_scene(scene_seed, n_boxes, n_cyls, 60.0);
    const uint32_t rays = beams * azimuth_steps;
    std::vector<float> tmp(static_cast<size_t>(rays) * 4);
    std::vector<uint8_t> ok(rays, 0);
    n_threads = std::max(1u, n_threads);
    auto work = [&](uint32_t tid) {
        const double pi = 3.14159265358979323846;
        for (uint32_t r = tid; r < rays; r += n_threads)
        {
            const uint32_t az_i = r / beams, beam = r % beams;
            const double elev = (beams > 1 ? elev_min_deg + (elev_max_deg - elev_min_deg) * beam / (beams - 1.0) : elev_min_deg) * pi / 180.0;
            const double az = 2.0 * pi * az_i / azimuth_steps;
            const double d[3] = {std::cos(elev) * std::cos(az), std::cos(elev) * std::sin(az), std::sin(elev)};
            const double o[3] = {sx, sy, 0.0};
            double t = cast(sc, o, d, 120.0);
            if (t > 120.0)
                continue;
            const uint64_t h1 = mix(noise_seed * 0x9E3779B97F4A7C15ull + r * 2ull);
            const uint64_t h2 = mix(noise_seed * 0x9E3779B97F4A7C15ull + r * 2ull + 1ull);
            const double g = std::sqrt(-2.0 * std::log(u01(h1))) * std::cos(2.0 * pi * u01(h2));
            t += 0.01 * g;
            tmp[static_cast<size_t>(r) * 4 + 0] = quant_mm(o[0] + t * d[0]);
            tmp[static_cast<size_t>(r) * 4 + 1] = quant_mm(o[1] + t * d[1]);
            tmp[static_cast<size_t>(r) * 4 + 2] = quant_mm(o[2] + t * d[2]);
            tmp[static_cast<size_t>(r) * 4 + 3] = static_cast<float>((mix(h1 ^ h2) % 100) / 100.0);
            ok[r] = 1;
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < n_threads; ++t)
        th.emplace_back(work, t);
    work(0);
    for (auto &t : th)
        t.join();
    uint32_t n = 0;
    for (uint32_t r = 0; r < rays; ++r)
        if (ok[r])
        {
            for (int k = 0; k < 4; ++k)
                out[static_cast<size_t>(n) * 4 + k] = tmp[static_cast<size_t>(r) * 4 + k];
            ++n;
        }
    return n;
}

// Dense stress structures for the union-find / replay: Gaussian blobs and lattice walls.
// Returns the number of points written (capacity must be >= n_blobs*blob_pts + walls points).
uint32_t synth_stress(uint64_t seed, uint32_t n_blobs, uint32_t blob_pts, double blob_sigma, uint32_t n_walls,
                      double wall_len, double wall_height, double lattice, float *out)
{
    SplitMix64 rng(seed ^ 0xD1B54A32D192ED03ull);
    const double pi = 3.14159265358979323846;
    uint32_t n = 0;
    for (uint32_t b = 0; b < n_blobs; ++b)
    {
        const double cx = rng.uniform(-100.0, 100.0), cy = rng.uniform(-100.0, 100.0), cz = rng.uniform(-1.0, 3.0);
        for (uint32_t i = 0; i < blob_pts; ++i)
        {
            const double r1 = std::sqrt(-2.0 * std::log(rng.uniform() + 1e-300)), a1 = 2.0 * pi * rng.uniform();
            const double r2 = std::sqrt(-2.0 * std::log(rng.uniform() + 1e-300)), a2 = 2.0 * pi * rng.uniform();
            out[static_cast<size_t>(n) * 4 + 0] = quant_mm(cx + blob_sigma * r1 * std::cos(a1));
            out[static_cast<size_t>(n) * 4 + 1] = quant_mm(cy + blob_sigma * r1 * std::sin(a1));
            out[static_cast<size_t>(n) * 4 + 2] = quant_mm(cz + blob_sigma * r2 * std::cos(a2));
            out[static_cast<size_t>(n) * 4 + 3] = 0.5f;
            ++n;
        }
    }
    for (uint32_t w = 0; w < n_walls; ++w)
    {
        const bool along_x = (w & 1u) == 0u;
        const double fixed = (w < 2 ? -1.0 : 1.0) * (70.0 + 10.0 * w);
        const uint32_t nu = static_cast<uint32_t>(wall_len / lattice), nv = static_cast<uint32_t>(wall_height / lattice);
        for (uint32_t u = 0; u < nu; ++u)
            for (uint32_t v = 0; v < nv; ++v)
            {
                const double a = -wall_len / 2 + u * lattice, z = -1.0 + v * lattice;
                out[static_cast<size_t>(n) * 4 + 0] = quant_mm(along_x ? a : fixed);
                out[static_cast<size_t>(n) * 4 + 1] = quant_mm(along_x ? fixed : a);
                out[static_cast<size_t>(n) * 4 + 2] = quant_mm(z);
                out[static_cast<size_t>(n) * 4 + 3] = 0.25f;
                ++n;
            }
    }
    return n;
}

} // extern "C"
