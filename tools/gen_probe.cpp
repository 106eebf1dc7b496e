// Probe (not product, not oracle): the reference BFS of clustering.cpp:69-124 restated GENERATION-synchronously.
// A FIFO generation = the entries pushed by the expansions of the previous generation. Inside one generation the
// expanded entries are the lexicographically-first independent set of the "within inner radius" conflict graph in
// queue order; removal times and pushes then follow in closed form. Prints per-component depth statistics and checks
// the labels against the sequential algorithm.
//   g++ -O2 -std=c++17 -o /tmp/gen_probe tools/gen_probe.cpp && /tmp/gen_probe pts.bin rank.bin
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <unordered_map>
#include <vector>

struct P { float x, y, z; };
static inline float d2f(const P &a, const P &b)
{
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return dx * dx + dy * dy + dz * dz;
}

int main(int argc, char **argv)
{
    FILE *f = fopen(argv[1], "rb");
    fseek(f, 0, SEEK_END);
    const size_t m = ftell(f) / 12;
    fseek(f, 0, SEEK_SET);
    std::vector<P> pts(m);
    if (fread(pts.data(), 12, m, f) != m) return 1;
    fclose(f);
    std::vector<uint32_t> rank(m);
    f = fopen(argv[2], "rb");
    if (fread(rank.data(), 4, m, f) != m) return 1;
    fclose(f);
    const float r2 = 0.18f;
    const float rin2 = static_cast<float>(std::pow(1.0 - 0.5, 2) * static_cast<double>(r2));
    const float cell = 1.001f * std::sqrt(r2);
    auto key = [&](int cx, int cy, int cz) { return (uint64_t(cx + (1 << 20)) << 42) | (uint64_t(cy + (1 << 20)) << 21) | uint64_t(cz + (1 << 20)); };
    std::unordered_map<uint64_t, std::vector<uint32_t>> grid;
    std::vector<int> cx(m), cy(m), cz(m);
    for (size_t i = 0; i < m; ++i) {
        cx[i] = (int)std::floor(pts[i].x / cell); cy[i] = (int)std::floor(pts[i].y / cell); cz[i] = (int)std::floor(pts[i].z / cell);
        grid[key(cx[i], cy[i], cz[i])].push_back(i);
    }
    // neighbour lists within r, sorted by k-d rank (= radius_search output order)
    std::vector<std::vector<uint32_t>> nb(m);
    for (size_t i = 0; i < m; ++i) {
        for (int a = -1; a <= 1; ++a) for (int b = -1; b <= 1; ++b) for (int c = -1; c <= 1; ++c) {
            auto it = grid.find(key(cx[i] + a, cy[i] + b, cz[i] + c));
            if (it == grid.end()) continue;
            for (uint32_t k : it->second) if (d2f(pts[i], pts[k]) <= r2) nb[i].push_back(k);
        }
        std::sort(nb[i].begin(), nb[i].end(), [&](uint32_t a, uint32_t b) { return rank[a] < rank[b]; });
    }
    // sequential reference
    std::vector<int> lab_seq(m, -1);
    std::vector<uint64_t> cnt_seq;
    {
        std::vector<char> removed(m, 0);
        int label = 0;
        for (uint32_t i = 0; i < m; ++i) {
            if (removed[i]) continue;
            std::deque<uint32_t> q{i};
            uint64_t cnt = 0;
            while (!q.empty()) {
                uint32_t j = q.front(); q.pop_front();
                if (removed[j]) continue;
                for (uint32_t k : nb[j]) {
                    if (removed[k]) continue;
                    lab_seq[k] = label; ++cnt;
                    if (d2f(pts[j], pts[k]) <= rin2) removed[k] = 1; else q.push_back(k);
                }
            }
            cnt_seq.push_back(cnt); ++label;
        }
    }
    // generation-synchronous restatement
    std::vector<int> lab_gen(m, -1);
    std::vector<uint64_t> cnt_gen;
    std::vector<char> removed(m, 0);
    const uint32_t INF = 0xFFFFFFFFu;
    std::vector<uint32_t> rho(m, INF), first_pusher(m, INF), qpos(m, INF);
    int label = 0;
    uint64_t tot_gens = 0, tot_q = 0, tot_e = 0, tot_mis_rounds = 0, tot_cand_q = 0, tot_cand_e = 0;
    struct Row { uint32_t seed; uint64_t members, gens, q, e, mis_rounds, max_q, max_mis; };
    std::vector<Row> rows;
    for (uint32_t i = 0; i < m; ++i) {
        if (removed[i]) continue;
        std::vector<uint32_t> Q{i};
        uint64_t cnt = 0;
        Row row{i, 0, 0, 0, 0, 0, 0, 0};
        while (!Q.empty()) {
            const size_t n = Q.size();
            for (size_t p = 0; p < n; ++p) qpos[Q[p]] = p;
            // lexicographically-first independent set by rounds
            std::vector<int> st(n, 0);  // 0 unresolved, 1 IN, 2 OUT
            size_t unresolved = n;
            uint64_t rounds = 0;
            while (unresolved) {
                ++rounds;
                std::vector<int> nst = st;
                for (size_t p = 0; p < n; ++p) {
                    if (st[p]) continue;
                    bool any_in = false, any_unres = false;
                    for (uint32_t k : nb[Q[p]]) {
                        if (qpos[k] == INF || qpos[k] >= p) continue;
                        if (d2f(pts[Q[p]], pts[k]) > rin2) continue;
                        if (st[qpos[k]] == 1) any_in = true;
                        else if (st[qpos[k]] == 0) any_unres = true;
                    }
                    if (any_in) nst[p] = 2; else if (!any_unres) nst[p] = 1;
                }
                for (size_t p = 0; p < n; ++p) if (!st[p] && nst[p]) --unresolved;
                st = nst;
            }
            // removal times of this generation
            std::vector<uint32_t> touched;
            for (size_t p = 0; p < n; ++p) if (st[p] == 1) {
                for (uint32_t k : nb[Q[p]]) if (!removed[k] && d2f(pts[Q[p]], pts[k]) <= rin2) {
                    if (rho[k] == INF) touched.push_back(k);
                    rho[k] = std::min<uint32_t>(rho[k], p);
                }
            }
            // touches and pushes
            std::vector<uint32_t> pushed;
            uint64_t ne = 0;
            for (size_t p = 0; p < n; ++p) if (st[p] == 1) {
                ++ne;
                tot_cand_e += nb[Q[p]].size();
                for (uint32_t k : nb[Q[p]]) {
                    if (removed[k] || rho[k] < p) continue;
                    lab_gen[k] = label; ++cnt;
                    if (d2f(pts[Q[p]], pts[k]) > rin2) {
                        if (first_pusher[k] == INF) pushed.push_back(k);
                        first_pusher[k] = std::min<uint32_t>(first_pusher[k], p);
                    }
                }
            }
            for (size_t p = 0; p < n; ++p) tot_cand_q += nb[Q[p]].size();
            for (size_t p = 0; p < n; ++p) qpos[Q[p]] = INF;
            for (uint32_t k : touched) { removed[k] = 1; rho[k] = INF; }
            // next generation: first occurrences, entries that died meanwhile dropped
            std::vector<uint32_t> next;
            for (uint32_t k : pushed) if (!removed[k]) next.push_back(k);
            std::sort(next.begin(), next.end(), [&](uint32_t a, uint32_t b) {
                return first_pusher[a] != first_pusher[b] ? first_pusher[a] < first_pusher[b] : rank[a] < rank[b]; });
            for (uint32_t k : pushed) first_pusher[k] = INF;
            row.gens++; row.q += n; row.e += ne; row.mis_rounds += rounds;
            row.max_q = std::max<uint64_t>(row.max_q, n); row.max_mis = std::max<uint64_t>(row.max_mis, rounds);
            Q.swap(next);
        }
        row.members = cnt;
        tot_gens += row.gens; tot_q += row.q; tot_e += row.e; tot_mis_rounds += row.mis_rounds;
        rows.push_back(row);
        cnt_gen.push_back(cnt); ++label;
    }

    // windowed restatement = what replay_gen.cuh does: the next <= W FIFO entries form a window; the expanded ones are
    // the lexicographically-first independent set; removed_before / queued_before of a candidate follow from geometry
    // against the earlier expanded entries of the window; pushes enter the FIFO by (window position, rank).
    std::vector<int> lab_win(m, -1);
    std::vector<uint64_t> cnt_win;
    {
        const uint32_t W = argc > 3 ? atoi(argv[3]) : 256;
        std::vector<char> rem(m, 0), queued(m, 0);
        int label = 0;
        uint64_t hist_w[5] = {0}, hist_n[5] = {0}, hist_in[5] = {0}; uint64_t tot_alive = 0, tot_T = 0, windows = 0, tot_n = 0, tot_in = 0, tot_push = 0, max_windows_seed = 0, tot_near = 0, tot_inner_tests = 0;
        for (uint32_t i = 0; i < m; ++i) {
            if (rem[i]) continue;
            std::vector<uint32_t> fifo{i};
            queued[i] = 1;
            size_t head = 0;
            uint64_t cnt = 0, wseed = 0;
            while (head < fifo.size()) {
                const size_t n = std::min<size_t>(W, fifo.size() - head);
                std::vector<uint32_t> win(fifo.begin() + head, fifo.begin() + head + n);
                head += n;
                ++windows; ++wseed; tot_n += n;
                for (size_t p = 0; p < n; ++p) tot_alive += !rem[win[p]];
                std::vector<int> in(n, 0);
                for (size_t p = 0; p < n; ++p) {
                    if (rem[win[p]]) continue;
                    bool out = false;
                    for (size_t e = 0; e < p && !out; ++e) if (in[e] && d2f(pts[win[e]], pts[win[p]]) <= rin2) out = true;
                    in[p] = !out;
                }
                std::vector<uint32_t> newly;
                std::vector<std::pair<uint64_t, uint32_t>> pushes;
                for (size_t p = 0; p < n; ++p) if (in[p]) {
                    ++tot_in;
                    const P &pj = pts[win[p]];
                    { const uint32_t q = win[p];
                      for (int a = -1; a <= 1; ++a) for (int b = -1; b <= 1; ++b) for (int c = -1; c <= 1; ++c) {
                          auto it = grid.find(key(cx[q] + a, cy[q] + b, cz[q] + c));
                          if (it != grid.end()) tot_T += it->second.size(); } }
                    std::vector<size_t> near;
                    for (size_t e = 0; e < p; ++e) if (in[e] && d2f(pts[win[e]], pj) <= 4.01f * r2) near.push_back(e);
                    tot_near += near.size();
                    for (uint32_t k : nb[win[p]]) {
                        if (rem[k]) continue;  // state at window start
                        bool removed_before = false, shared = false;
                        for (size_t e : near) {
                            const float dj = d2f(pts[win[e]], pts[k]);
                            removed_before |= dj <= rin2; shared |= dj <= r2; ++tot_inner_tests;
                        }
                        if (removed_before) continue;
                        lab_win[k] = label; ++cnt;
                        if (d2f(pj, pts[k]) <= rin2) newly.push_back(k);
                        else if (!queued[k] && !shared) { pushes.push_back({(uint64_t(p) << 32) | rank[k], k}); }
                    }
                }
                { size_t nin = 0; for (size_t p = 0; p < n; ++p) nin += in[p];
                  int bin = n <= 1 ? 0 : n <= 8 ? 1 : n <= 32 ? 2 : n <= 128 ? 3 : 4;
                  hist_w[bin]++; hist_n[bin] += n; hist_in[bin] += nin; }
                for (uint32_t k : newly) rem[k] = 1;
                std::sort(pushes.begin(), pushes.end());
                for (auto &pr : pushes) { queued[pr.second] = 1; fifo.push_back(pr.second); }
                tot_push += pushes.size();
            }
            max_windows_seed = std::max(max_windows_seed, wseed);
            cnt_win.push_back(cnt); ++label;
        }
        const char *names[5] = {"n=1", "n<=8", "n<=32", "n<=128", "n<=256"};
        for (int b = 0; b < 5; ++b) printf("  windows %-7s: %6llu windows, %7llu entries, %6llu expanded\n", names[b], (unsigned long long)hist_w[b], (unsigned long long)hist_n[b], (unsigned long long)hist_in[b]);
        size_t badw = 0, badcw = cnt_win.size() != cnt_seq.size();
        for (size_t i = 0; i < m; ++i) badw += lab_seq[i] != lab_win[i];
        for (size_t i = 0; i < std::min(cnt_seq.size(), cnt_win.size()); ++i) badcw += cnt_seq[i] != cnt_win[i];
        printf("window model W=%u: label mismatches %zu count mismatches %zu; windows %llu (longest seed %llu) entries %llu expanded %llu pushes %llu near/expanded %.2f inner tests %llu alive entries %llu cell candidates/expanded %.1f\n",
               W, badw, badcw, (unsigned long long)windows, (unsigned long long)max_windows_seed, (unsigned long long)tot_n,
               (unsigned long long)tot_in, (unsigned long long)tot_push, double(tot_near) / double(std::max<uint64_t>(1, tot_in)), (unsigned long long)tot_inner_tests, (unsigned long long)tot_alive, double(tot_T) / double(std::max<uint64_t>(1, tot_in)));
    }
    size_t bad = 0;
    for (size_t i = 0; i < m; ++i) bad += lab_seq[i] != lab_gen[i];
    size_t badc = cnt_seq.size() != cnt_gen.size();
    for (size_t i = 0; i < std::min(cnt_seq.size(), cnt_gen.size()); ++i) badc += cnt_seq[i] != cnt_gen[i];
    printf("points %zu seeds %zu label mismatches %zu count mismatches %zu\n", m, cnt_seq.size(), bad, badc);
    printf("total: gens %llu queue entries %llu expansions %llu mis rounds %llu cand(queue) %llu cand(expanded) %llu\n",
           (unsigned long long)tot_gens, (unsigned long long)tot_q, (unsigned long long)tot_e, (unsigned long long)tot_mis_rounds,
           (unsigned long long)tot_cand_q, (unsigned long long)tot_cand_e);
    std::sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.e > b.e; });
    printf("top seeds by expansions: seed touches gens queue exp mis_rounds max_q max_mis\n");
    for (size_t i = 0; i < std::min<size_t>(rows.size(), 12); ++i)
        printf("  %u %llu %llu %llu %llu %llu %llu %llu\n", rows[i].seed, (unsigned long long)rows[i].members, (unsigned long long)rows[i].gens,
               (unsigned long long)rows[i].q, (unsigned long long)rows[i].e, (unsigned long long)rows[i].mis_rounds,
               (unsigned long long)rows[i].max_q, (unsigned long long)rows[i].max_mis);
    return 0;
}
