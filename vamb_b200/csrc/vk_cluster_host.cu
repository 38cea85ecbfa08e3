// Native host driver of the medoid clusterer: the decision logic of vamb/cluster.py
// (ClusterGenerator.__next__ :298-316, find_cluster :545-604, wander_medoid :415-450,
// find_threshold :452-543, get_next_seed :342-384, update_successes :386-413) in C++, so that one
// emitted cluster costs one foreign call instead of ~10 Python round trips.  The device work is the
// same set of kernels (vk_probe_sync, vk_eval_candidates_sync, vk_select_members_sync,
// vk_compact_rows_sync); vamb_b200/cluster.py keeps the line-by-line Python rendition of this logic
// and the two are tested to emit identical clusters.
//
// The reference samples medoid candidates with Python's random.Random(rng_seed).sample(); to emit the
// same clusters this file restates CPython's generator: MT19937 seeded by init_by_array over the
// 32-bit words of |seed|, getrandbits(k) = genrand_uint32() >> (32 - k), _randbelow by rejection,
// and sample() with its pool / set-rejection switch at n <= 21 + 4^ceil(log4(3k)).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <deque>
#include <new>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "vk_common.cuh"

namespace {

struct MT19937 {
    uint32_t mt[624];
    int idx;
    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void init_by_array(const uint32_t *key, int len) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = (624 > len ? 624 : len); k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            ++i; ++j;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
            if (j >= len) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            ++i;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    // random.Random._randbelow_with_getrandbits for 0 < n < 2^32
    uint32_t randbelow(uint32_t n) {
        int k = 0;
        for (uint32_t t = n; t; t >>= 1) ++k;
        uint32_t r = next() >> (32 - k);
        while (r >= n) r = next() >> (32 - k);
        return r;
    }
    // random.Random.sample(population, k)
    void sample(const std::vector<int32_t> &population, int k, std::vector<int32_t> &result) {
        const int n = (int)population.size();
        result.assign(k, 0);
        double setsize = 21.0;
        if (k > 5) setsize += pow(4.0, ceil(log((double)k * 3.0) / log(4.0)));
        if ((double)n <= setsize) {
            std::vector<int32_t> pool(population);
            for (int i = 0; i < k; ++i) {
                const uint32_t j = randbelow((uint32_t)(n - i));
                result[i] = pool[j];
                pool[j] = pool[n - i - 1];
            }
        } else {
            std::unordered_set<uint32_t> selected;
            for (int i = 0; i < k; ++i) {
                uint32_t j = randbelow((uint32_t)n);
                while (selected.count(j)) j = randbelow((uint32_t)n);
                selected.insert(j);
                result[i] = population[j];
            }
        }
    }
};

struct Probe {
    int32_t medoid;
    unsigned __int128 density;
    uint64_t hist[VK_NBINS];
    int32_t n_within, n_lt, n_nl, rank;
    std::vector<int32_t> within;  // ascending rows
};

struct State {
    vk_cluster_config c;
    int cur;  // which of the two buffer sets is live
    int64_t n_act;
    std::vector<int32_t> indices;    // original id of every live device row (ascending)
    std::vector<uint8_t> kept_host;  // host mirror of the device mask
    std::vector<int64_t> order;
    int64_t order_index;
    double pvr;
    std::deque<uint8_t> attempts;
    int successes;
    int64_t n_emitted, n_remaining;
    MT19937 rng;
    float pdf[31];
    float edges[VK_NBINS + 1];
    // mapped completion (vk_probe_mapped / vk_eval_candidates_mapped): pinned results + flags, device tickets
    vk_probe_header *hdr_pin = nullptr;
    uint64_t *cand_pin = nullptr;
    int32_t *flags_pin = nullptr;   // [0] probe, [1] candidate evaluation
    int32_t *tickets_dev = nullptr; // [0] probe, [1] candidate evaluation, [2] work counter of the probe
    int32_t seq = 0;
    // lazy medoid moves (vk_eval_candidates_lists): device accumulators, pinned results and id lists
    uint64_t *cand2_dev = nullptr, *cand2_pin = nullptr;
    int32_t *within_pin = nullptr, *within_dev = nullptr;
    int64_t n_moves_lazy = 0, n_rebases = 0, sum_nnl = 0;  // sum_nnl: neighbour-list sizes over all candidate evaluations
    bool lazy_enabled = true;  // VAMB_B200_CLUSTER_LAZY=0: a full scan per move, as round 1 (same clusters)
    std::vector<int64_t> members;
    // per-wander sets over device rows as stamp arrays (row -> id of the wander that set it): `tried`, and the cache of
    // candidate evaluations (stamp + slot in cache_vec).  Hash containers cost ~1 s of the 6.3 s C2 clustering.
    std::vector<uint32_t> tried_stamp, cache_stamp;
    std::vector<int32_t> cache_slot;
    uint32_t wander_id = 0;
    int64_t n_probes, n_evals, n_packs;
    double t_probe = 0.0, t_eval = 0.0, t_select = 0.0, t_pack = 0.0, t_total = 0.0;  // host wall seconds per call kind

    float *M() const { return cur ? c.matrix2 : c.matrix; }
    float *LEN() const { return cur ? c.lengths2 : c.lengths; }
    uint8_t *KEPT() const { return cur ? c.kept2 : c.kept; }
    int32_t *ORIG() const { return cur ? c.orig2 : c.orig; }
};

struct Stopwatch {  // adds the enclosing scope's wall time to `acc`
    double &acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit Stopwatch(double &a) : acc(a) {}
    ~Stopwatch() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

int do_probe(State &st, int32_t row, Probe &p) {
    Stopwatch sw(st.t_probe);
    ++st.n_probes;
    const vk_cluster_config &c = st.c;
    if (vk_probe_mapped(st.M(), st.LEN(), st.KEPT(), st.n_act, c.d, row, c.nl_radius, c.edges, c.hdr, c.within_overflow,
                        c.nl_rows, c.nl_dists, st.hdr_pin, st.tickets_dev, st.flags_pin, ++st.seq, st.tickets_dev + 2, c.stream))
        return 1;
    const vk_probe_header *h = st.hdr_pin;
    p.medoid = row;
    p.density = ((unsigned __int128)h->density_hi << 12) + h->density_lo;
    memcpy(p.hist, h->hist, sizeof(p.hist));
    p.n_within = h->n_within; p.n_lt = h->n_lt; p.n_nl = h->n_nl; p.rank = h->rank;
    const int inl = p.n_within < VK_PROBE_INLINE ? p.n_within : VK_PROBE_INLINE;
    p.within.assign(h->within, h->within + inl);
    if (p.n_within > VK_PROBE_INLINE) {
        const size_t extra = (size_t)(p.n_within - VK_PROBE_INLINE);
        p.within.resize((size_t)p.n_within);
        VK_CUDA(cudaMemcpyAsync(p.within.data() + VK_PROBE_INLINE, c.within_overflow + VK_PROBE_INLINE,
                                extra * sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)c.stream));
        VK_CUDA(cudaStreamSynchronize((cudaStream_t)c.stream));
    }
    std::sort(p.within.begin(), p.within.end());
    return 0;
}

int do_eval(State &st, const Probe &p, const std::vector<int32_t> &rows, std::vector<unsigned __int128> &dens) {
    Stopwatch sw(st.t_eval);
    const vk_cluster_config &c = st.c;
    dens.clear();
    for (size_t i = 0; i < rows.size(); i += VK_MAX_CAND) {
        const int n = (int)std::min<size_t>(VK_MAX_CAND, rows.size() - i);
        ++st.n_evals;
        if (vk_eval_candidates_mapped(st.M(), st.LEN(), c.d, c.nl_rows, c.nl_dists, p.n_nl, c.prune_radius,
                                      rows.data() + i, n, c.cand_out, st.cand_pin, st.tickets_dev + 1, st.flags_pin + 1,
                                      ++st.seq, c.stream))
            return 1;
        for (int k = 0; k < n; ++k)
            dens.push_back(((unsigned __int128)st.cand_pin[VK_MAX_CAND + k] << 12) + st.cand_pin[k]);
    }
    return 0;
}

int do_select(State &st, const Probe &p, float threshold) {
    Stopwatch sw(st.t_select);
    const vk_cluster_config &c = st.c;
    if (vk_select_members_sync(c.nl_rows, c.nl_dists, p.n_nl, threshold, st.ORIG(), st.KEPT(), c.members,
                               c.members_host, c.members_host_cap, c.stream))
        return 1;
    const int cnt = c.members_host[0];
    st.members.resize((size_t)cnt);
    if (cnt + 1 <= c.members_host_cap) {
        for (int i = 0; i < cnt; ++i) st.members[i] = c.members_host[1 + i];
    } else {
        std::vector<int32_t> tmp((size_t)cnt);
        VK_CUDA(cudaMemcpyAsync(tmp.data(), c.members + 1, sizeof(int32_t) * (size_t)cnt, cudaMemcpyDeviceToHost,
                                (cudaStream_t)c.stream));
        VK_CUDA(cudaStreamSynchronize((cudaStream_t)c.stream));
        for (int i = 0; i < cnt; ++i) st.members[i] = tmp[i];
    }
    std::sort(st.members.begin(), st.members.end());
    return 0;
}

int do_pack(State &st) {
    Stopwatch sw(st.t_pack);
    const vk_cluster_config &c = st.c;
    ++st.n_packs;
    int64_t n_out = 0;
    const int nxt = st.cur ^ 1;
    if (vk_compact_rows_sync(st.M(), st.LEN(), st.ORIG(), st.KEPT(), st.n_act, c.d, nxt ? c.matrix2 : c.matrix,
                             nxt ? c.lengths2 : c.lengths, nxt ? c.orig2 : c.orig, nxt ? c.kept2 : c.kept,
                             c.tile_scratch, &n_out, c.stream))
        return 1;
    st.cur = nxt;
    size_t w = 0;
    for (size_t i = 0; i < (size_t)st.n_act; ++i)
        if (st.kept_host[i]) st.indices[w++] = st.indices[i];
    st.indices.resize(w);
    st.kept_host.assign(w, 1);
    if ((int64_t)w != n_out) {
        vk_set_error("vk_cluster: host/device live-row counts disagree (%lld vs %lld)", (long long)w, (long long)n_out);
        return 1;
    }
    st.n_act = n_out;
    return 0;
}

// vamb/cluster.py:342-384
int32_t next_seed(State &st) {
    int64_t n_orig = (int64_t)st.order.size();
    int64_t i = st.order_index - 1;
    for (;;) {
        i = (i + 1) % n_orig;
        if (i == 0 && st.n_emitted > 0) {
            size_t w = 0;
            for (size_t k = 0; k < st.order.size(); ++k)
                if (st.order[k] > -1) st.order[w++] = st.order[k];
            st.order.resize(w);
            n_orig = (int64_t)w;
        }
        const int64_t o = st.order[(size_t)i];
        if (o == -1) continue;
        const auto it = std::lower_bound(st.indices.begin(), st.indices.end(), (int32_t)o);
        const size_t row = (size_t)(it - st.indices.begin());
        if (it == st.indices.end() || *it != (int32_t)o || !st.kept_host[row]) {
            st.order[(size_t)i] = -1;
            continue;
        }
        st.order_index = i + 1;
        return (int32_t)row;
    }
}

// vamb/cluster.py:386-413
void update_successes(State &st, bool success) {
    if ((int)st.attempts.size() == st.c.windowsize) {
        st.successes -= st.attempts.front();
        st.attempts.pop_front();
    }
    st.successes += success ? 1 : 0;
    st.attempts.push_back(success ? 1 : 0);
    if ((int)st.attempts.size() == st.c.windowsize && st.successes < st.c.minsuccesses) {
        st.pvr += 0.1;
        st.attempts.clear();
        st.successes = 0;
        st.order_index = 0;
    }
}

constexpr int WITHIN_CAP = 1024;     // ids per candidate in the pinned list buffer
constexpr float R_EVAL = 0.1199f;    // d(candidate, base) up to which the base's 0.3-neighbour list covers the
                                     // candidate's whole 0.05-neighbourhood: acos(1 - 2 * 0.1203) + acos(0.9) = acos(0.4)

struct EvalLists {
    std::vector<unsigned __int128> dens;
    std::vector<int64_t> cnt;
    std::vector<float> dbase;
};

int do_eval_lists(State &st, const Probe &base, float prune, const std::vector<int32_t> &rows, unsigned __int128 min_density,
                  EvalLists &out) {
    Stopwatch sw(st.t_eval);
    const vk_cluster_config &c = st.c;
    const int n = (int)rows.size();
    ++st.n_evals;
    st.sum_nnl += base.n_nl;
    if (vk_eval_candidates_lists(st.M(), st.LEN(), c.d, c.nl_rows, c.nl_dists, base.n_nl, prune, rows.data(), n, base.medoid,
                                 (uint64_t)(min_density >> 12), (uint64_t)(min_density & 4095), st.cand2_dev, st.cand2_pin, st.within_dev, st.within_pin, WITHIN_CAP, st.tickets_dev + 1,
                                 st.flags_pin + 1, ++st.seq, c.stream))
        return 1;
    out.dens.resize((size_t)n);
    out.cnt.resize((size_t)n);
    out.dbase.resize((size_t)n);
    for (int k = 0; k < n; ++k) {
        out.dens[(size_t)k] = ((unsigned __int128)st.cand2_pin[VK_LIST_CAND + k] << 12) + st.cand2_pin[k];
        out.cnt[(size_t)k] = (int64_t)st.cand2_pin[2 * VK_LIST_CAND + k];
        const uint32_t bits = (uint32_t)st.cand2_pin[3 * VK_LIST_CAND + k];
        memcpy(&out.dbase[(size_t)k], &bits, sizeof(float));
    }
    return 0;
}

// vamb/cluster.py:415-450.  The reference calls sample_medoid (a full scan) for every candidate and for every
// move.  Here one full scan ("probe") at a BASE medoid leaves its neighbour list (rows within 0.3) on the device; the
// candidates of a round are evaluated together over that list, and a move to a winning candidate needs no scan at all
// as long as the list still covers the new candidates' 0.05-neighbourhoods (d(candidate, base) <= R_EVAL): the
// winner's within-set and density come out of the evaluation.  Only when a sampled candidate lies too far from the base
// is the current medoid probed (it becomes the base) and the SAME sample evaluated again; the final medoid is probed
// once for its histogram / neighbour list.  Densities are pure functions of (medoid, live rows), so the decisions --
// and the Python RNG call sequence -- are exactly the reference's.
int wander(State &st, int32_t seed, Probe &probe, int32_t &seed_rank) {
    if (st.tried_stamp.size() < (size_t)st.n_act) {  // rows only ever renumber downwards (compaction)
        st.tried_stamp.assign((size_t)st.n_act, 0u);
        st.cache_stamp.assign((size_t)st.n_act, 0u);
        st.cache_slot.assign((size_t)st.n_act, 0);
    }
    if (++st.wander_id == 0u) {  // wrapped: stale stamps could alias
        std::fill(st.tried_stamp.begin(), st.tried_stamp.end(), 0u);
        std::fill(st.cache_stamp.begin(), st.cache_stamp.end(), 0u);
        st.wander_id = 1u;
    }
    const uint32_t wid = st.wander_id;
    auto is_tried = [&](int32_t r) { return st.tried_stamp[(size_t)r] == wid; };
    auto set_tried = [&](int32_t r) { st.tried_stamp[(size_t)r] = wid; };
    set_tried(seed);
    if (do_probe(st, seed, probe)) return 1;
    seed_rank = probe.rank;
    const bool list_is_everything = !(st.c.nl_radius < 1e30f);
    const bool lazy = st.lazy_enabled && st.c.maxsteps <= VK_LIST_CAND && st.cand2_dev != nullptr &&
                      (list_is_everything || st.c.nl_radius == 0.3f);
    unsigned __int128 local = probe.density;
    int32_t cur = seed;                         // current medoid; probe describes the base (probe.medoid)
    std::vector<int32_t> cur_within = probe.within;
    std::vector<int32_t> cand, sampled;
    std::vector<unsigned __int128> dens;
    EvalLists ev;
    struct Cached {
        unsigned __int128 dens;
        std::vector<int32_t> within;
        bool lists_ok;
        bool sorted;  // `within` is sorted only when it becomes the current within-set (one winner per round, not 64 lists)
    };
    std::vector<Cached> cache_vec;  // what the candidate evaluations of this wander returned, by slot
    auto cached = [&](int32_t r) -> Cached * {
        return st.cache_stamp[(size_t)r] == wid ? &cache_vec[(size_t)st.cache_slot[(size_t)r]] : nullptr;
    };
    auto cache_new = [&](int32_t r) -> Cached & {
        if (Cached *c = cached(r)) return *c;
        st.cache_stamp[(size_t)r] = wid;
        st.cache_slot[(size_t)r] = (int32_t)cache_vec.size();
        cache_vec.emplace_back();
        return cache_vec.back();
    };
    std::vector<int32_t> batch;
    auto rebase = [&]() -> int {                // full scan at the current medoid
        const unsigned __int128 want = local;
        if (do_probe(st, cur, probe)) return 1;
        ++st.n_rebases;
        if (probe.density != want || probe.within != cur_within) {
            vk_set_error("vk_cluster: probe and candidate evaluation disagree");
            return 1;
        }
        return 0;
    };
    for (;;) {
        cand.clear();
        for (int32_t r : cur_within)
            if (!is_tried(r)) cand.push_back(r);
        const int k = (int)std::min<size_t>(cand.size(), (size_t)st.c.maxsteps);
        st.rng.sample(cand, k, sampled);
        if (sampled.empty()) break;
        int winner = -1;
        if (!lazy) {
            if (cur != probe.medoid && rebase()) return 1;
            if (do_eval(st, probe, sampled, dens)) return 1;
            for (size_t i = 0; i < sampled.size(); ++i) {
                set_tried(sampled[i]);
                if (dens[i] > local) { winner = (int)i; break; }
            }
            if (winner < 0) break;
            const unsigned __int128 want = dens[(size_t)winner];
            if (do_probe(st, sampled[(size_t)winner], probe)) return 1;
            if (probe.density != want) {
                vk_set_error("vk_cluster: probe and candidate densities disagree");
                return 1;
            }
            cur = sampled[(size_t)winner];
            cur_within = probe.within;
            local = probe.density;
            continue;
        }
        // Densities (and within-sets) are pure functions of (row, live rows), and no row changes during one wander: what
        // a launch computes stays valid for the rest of this wander.  Every launch therefore evaluates the sampled
        // candidates that are not known yet PLUS as many further members of the current within-set as fit (the next
        // rounds sample from exactly those), so most rounds -- and most moves between overlapping within-sets -- need
        // no device call at all.  Only sampled candidates enter `tried`, in the reference's order.
        for (;;) {
            batch.clear();
            for (int32_t c : sampled)
                if (!cached(c)) batch.push_back(c);
            if (!batch.empty()) {
                const size_t n_need = batch.size();
                for (int32_t c : cur_within) {
                    if (batch.size() >= (size_t)VK_LIST_CAND) break;
                    if (!is_tried(c) && !cached(c) && std::find(batch.begin(), batch.end(), c) == batch.end())
                        batch.push_back(c);
                }
                const bool at_base = cur == probe.medoid;
                // candidates of the base itself lie within 0.05 of it: rows near them are within 0.19 (prune radius)
                // id lists come back only for candidates denser than the current medoid: `local` only grows during a
                // wander, so no other candidate can ever be moved to
                if (do_eval_lists(st, probe, at_base ? st.c.prune_radius : st.c.nl_radius, batch, local, ev)) return 1;
                for (size_t i = 0; i < batch.size(); ++i) {
                    const bool ok = at_base || list_is_everything || ev.dbase[i] <= R_EVAL;
                    if (!ok) continue;  // the list may not cover this candidate's neighbourhood: not usable
                    Cached &cc = cache_new(batch[i]);
                    cc.dens = ev.dens[i];
                    cc.lists_ok = ev.cnt[i] <= WITHIN_CAP && ev.dens[i] > local;  // else: never a move target (or truncated)
                    cc.sorted = false;
                    if (cc.lists_ok) {
                        const int32_t *ids = st.within_pin + i * WITHIN_CAP;
                        cc.within.assign(ids, ids + ev.cnt[i]);
                    }
                }
                (void)n_need;
            }
            bool need_rebase = false;
            winner = -1;
            for (size_t i = 0; i < sampled.size(); ++i) {
                const Cached *it = cached(sampled[i]);
                if (!it) { need_rebase = true; break; }  // beyond the coverage radius of the base
                set_tried(sampled[i]);
                if (it->dens > local) { winner = (int)i; break; }
            }
            if (!need_rebase) break;
            if (rebase()) return 1;  // then evaluate what is still unknown of the same sample against the new base
        }
        if (winner < 0) break;
        const size_t w = (size_t)winner;
        Cached &cw = *cached(sampled[w]);
        if (!cw.lists_ok) {  // id list truncated: take the winner's within-set from a full scan
            const unsigned __int128 want = cw.dens;
            if (do_probe(st, sampled[w], probe)) return 1;
            if (probe.density != want) {
                vk_set_error("vk_cluster: probe and candidate densities disagree");
                return 1;
            }
            cur_within = probe.within;
        } else {
            if (!cw.sorted) {  // ascending row order = the reference's candidate order (vamb/cluster.py:626)
                std::sort(cw.within.begin(), cw.within.end());
                cw.sorted = true;
            }
            cur_within = cw.within;
            ++st.n_moves_lazy;
        }
        cur = sampled[w];
        local = cw.dens;
    }
    if (cur != probe.medoid && rebase()) return 1;  // the final medoid's histogram / neighbour list / loner count
    return 0;
}

// vamb/cluster.py:452-543.  Returns 0 = loner, 1 = no threshold, 2 = (threshold, observed_pvr)
int find_threshold(State &st, const Probe &p, double &threshold, double &observed_pvr) {
    if (p.n_lt == 1) return 0;
    float hist[VK_NBINS];
    for (int i = 0; i < VK_NBINS; ++i) hist[i] = (float)p.hist[i];  // exact sums, rounded once
    float dens[VK_NBINS + 30];
    for (int i = 0; i < VK_NBINS + 30; ++i) dens[i] = 0.0f;
    for (int i = 0; i < VK_NBINS; ++i)
        for (int j = 0; j < 31; ++j) {
            const float prod = st.pdf[j] * hist[i];  // two roundings, as in the reference (no fma)
            dens[i + j] = dens[i + j] + prod;
        }
    double peak_density = 0.0, minimum_x = 0.0, density_at_minimum = 0.0, x = 0.0;
    bool peak_over = false, have_threshold = false;
    const double delta_x = 0.3 / (double)VK_NBINS;
    for (int i = 0; i < VK_NBINS; ++i) {
        const double density = (double)dens[15 + i];
        if (!peak_over && density > peak_density) {
            if (x > 0.1) return 1;
            peak_density = density;
        }
        if (!peak_over && density < 0.6 * peak_density) {
            peak_over = true;
            density_at_minimum = density;
        }
        if (peak_over && density > 1.5 * density_at_minimum) break;
        if (peak_over && density < density_at_minimum) {
            minimum_x = x;
            density_at_minimum = density;
            if (density < st.pvr * peak_density) {
                threshold = minimum_x;
                have_threshold = true;
            }
        }
        x += delta_x;
    }
    if (!have_threshold) return 1;
    if (threshold > 0.2 + st.pvr) return 1;
    observed_pvr = density_at_minimum / peak_density;
    return 2;
}

}  // namespace

extern "C" void vk_cluster_destroy(void *handle);

extern "C" int vk_cluster_create(void **handle, const vk_cluster_config *cfg) {
    State *st = new (std::nothrow) State();
    if (!st) {
        vk_set_error("vk_cluster_create: out of memory");
        return 1;
    }
    st->c = *cfg;
    if (const char *v = getenv("VAMB_B200_CLUSTER_LAZY")) st->lazy_enabled = atoi(v) != 0;
    st->cur = 0;
    st->n_act = cfg->n;
    st->indices.resize((size_t)cfg->n);
    for (int64_t i = 0; i < cfg->n; ++i) st->indices[(size_t)i] = (int32_t)i;
    st->kept_host.assign((size_t)cfg->n, 1);
    st->order.assign(cfg->order_host, cfg->order_host + cfg->n);
    st->order_index = 0;
    st->pvr = 0.1;
    st->successes = 0;
    st->n_emitted = 0;
    st->n_remaining = cfg->n;
    st->n_probes = st->n_evals = st->n_packs = 0;
    st->rng.init_by_array(cfg->seed_key, cfg->seed_key_len);
    memcpy(st->pdf, cfg->normalpdf_host, sizeof(st->pdf));
    // mapped completion: pinned (device-visible) result buffers, and accumulators that start out zeroed
    cudaStream_t s = (cudaStream_t)cfg->stream;
    if (cudaHostAlloc((void **)&st->hdr_pin, sizeof(vk_probe_header), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostAlloc((void **)&st->cand_pin, sizeof(uint64_t) * 3 * VK_MAX_CAND, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostAlloc((void **)&st->flags_pin, sizeof(int32_t) * 2, cudaHostAllocMapped) != cudaSuccess ||
        cudaMalloc((void **)&st->tickets_dev, sizeof(int32_t) * 4) != cudaSuccess ||
        cudaHostAlloc((void **)&st->cand2_pin, sizeof(uint64_t) * 4 * VK_LIST_CAND, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostAlloc((void **)&st->within_pin, sizeof(int32_t) * VK_LIST_CAND * WITHIN_CAP, cudaHostAllocMapped) != cudaSuccess ||
        cudaMalloc((void **)&st->cand2_dev, sizeof(uint64_t) * VK_EVAL_SCRATCH_U64) != cudaSuccess ||
        cudaMalloc((void **)&st->within_dev, sizeof(int32_t) * VK_EVAL_SUBS * VK_LIST_CAND * WITHIN_CAP) != cudaSuccess ||
        cudaMemsetAsync(st->cand2_dev, 0, sizeof(uint64_t) * VK_EVAL_SCRATCH_U64, s) != cudaSuccess ||
        cudaMemsetAsync(st->tickets_dev, 0, sizeof(int32_t) * 4, s) != cudaSuccess ||
        cudaMemsetAsync(cfg->hdr, 0, sizeof(vk_probe_header), s) != cudaSuccess ||
        cudaMemsetAsync(cfg->cand_out, 0, sizeof(uint64_t) * 3 * VK_MAX_CAND, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        vk_set_error("vk_cluster_create: %s", cudaGetErrorString(cudaGetLastError()));
        vk_cluster_destroy(st);
        return 1;
    }
    memset(st->hdr_pin, 0, sizeof(vk_probe_header));
    st->flags_pin[0] = st->flags_pin[1] = 0;
    *handle = st;
    return 0;
}

extern "C" void vk_cluster_destroy(void *handle) {
    State *st = static_cast<State *>(handle);
    if (!st) return;
    if (st->hdr_pin) cudaFreeHost(st->hdr_pin);
    if (st->cand_pin) cudaFreeHost(st->cand_pin);
    if (st->flags_pin) cudaFreeHost(st->flags_pin);
    if (st->tickets_dev) cudaFree(st->tickets_dev);
    if (st->cand2_pin) cudaFreeHost(st->cand2_pin);
    if (st->within_pin) cudaFreeHost(st->within_pin);
    if (st->cand2_dev) cudaFree(st->cand2_dev);
    if (st->within_dev) cudaFree(st->within_dev);
    delete st;
}

extern "C" int vk_cluster_stats(void *handle, int64_t *out8) {
    State *st = static_cast<State *>(handle);
    out8[0] = st->n_probes; out8[1] = st->n_evals; out8[2] = st->n_packs; out8[3] = st->n_act;
    out8[4] = st->cur; out8[5] = st->successes; out8[6] = (int64_t)st->attempts.size(); out8[7] = st->order_index;
    return 0;
}

// host wall-clock seconds spent so far in: probes, candidate evaluations, member selections, packs, all of vk_cluster_next;
// then the number of medoid moves made without a scan and the number of re-basing probes
extern "C" int vk_cluster_timing(void *handle, double *out5 /* 8 doubles */) {
    State *st = static_cast<State *>(handle);
    out5[0] = st->t_probe; out5[1] = st->t_eval; out5[2] = st->t_select; out5[3] = st->t_pack; out5[4] = st->t_total;
    out5[5] = (double)st->n_moves_lazy; out5[6] = (double)st->n_rebases; out5[7] = (double)st->sum_nnl;
    return 0;
}

// One ClusterGenerator.__next__ (vamb/cluster.py:298-316).  Returns 0 = a cluster, 2 = exhausted, 1 = error.
extern "C" int vk_cluster_next(void *handle, vk_cluster_result *out) {
    State &st = *static_cast<State *>(handle);
    if (st.n_remaining == 0) return 2;
    Stopwatch sw_total(st.t_total);
    Probe probe;
    for (;;) {  // find_cluster, vamb/cluster.py:545-604
        const int32_t seed = next_seed(st);
        int32_t seed_rank = 0;
        if (wander(st, seed, probe, seed_rank)) return 1;
        double threshold = 0.0, observed = 0.0;
        const int kind = find_threshold(st, probe, threshold, observed);
        out->medoid = st.indices[(size_t)probe.medoid];
        out->seed = seed_rank;
        out->maximal_pvr = st.pvr;
        out->successes = st.successes;
        out->attempts = (int32_t)st.attempts.size();
        if (kind == 0) {
            if (do_select(st, probe, 0.0f)) return 1;
            if (st.members.size() != 1 || st.members[0] != out->medoid) {
                vk_set_error("vk_cluster: loner selection returned %zu members", st.members.size());
                return 1;
            }
            out->kind = 0; out->radius = NAN; out->observed_pvr = NAN;
            break;
        }
        if (kind == 1) {
            if (st.pvr > 0.55) {
                if (do_select(st, probe, (float)0.06)) return 1;
                out->kind = 1; out->radius = 0.06; out->observed_pvr = NAN;
                break;
            }
            update_successes(st, false);
            continue;
        }
        if (do_select(st, probe, (float)threshold)) return 1;
        out->kind = 2; out->radius = threshold; out->observed_pvr = observed;
        if (st.pvr < 0.55) update_successes(st, true);
        break;
    }
    st.n_emitted += 1;
    st.n_remaining -= (int64_t)st.members.size();
    for (int64_t m : st.members) {
        const auto it = std::lower_bound(st.indices.begin(), st.indices.end(), (int32_t)m);
        st.kept_host[(size_t)(it - st.indices.begin())] = 0;
    }
    out->members_host = st.members.data();
    out->n_members = (int64_t)st.members.size();
    out->peak_valley_ratio = st.pvr;
    out->n_remaining = st.n_remaining;
    if (st.n_remaining && (double)st.n_remaining < st.c.pack_fraction * (double)st.n_act)
        if (do_pack(st)) return 1;
    return 0;
}

extern "C" int vk_cluster_next_block(void *handle, vk_cluster_block *blk) {
    blk->n_clusters = 0;
    blk->n_members_total = 0;
    State &st = *static_cast<State *>(handle);
    vk_cluster_result r;
    while (blk->n_clusters < blk->max_clusters) {
        const int rc = vk_cluster_next(handle, &r);
        if (rc == 2) break;
        if (rc) return 1;
        const int64_t i = blk->n_clusters++;
        blk->medoid[i] = r.medoid; blk->seed[i] = r.seed; blk->n_members[i] = r.n_members;
        blk->maximal_pvr[i] = r.maximal_pvr; blk->observed_pvr[i] = r.observed_pvr; blk->radius[i] = r.radius;
        blk->kind[i] = r.kind; blk->successes[i] = r.successes; blk->attempts[i] = r.attempts;
        memcpy(blk->members + blk->n_members_total, r.members_host, sizeof(int64_t) * (size_t)r.n_members);
        blk->n_members_total += r.n_members;
    }
    blk->n_remaining = st.n_remaining;
    blk->peak_valley_ratio = st.pvr;
    return 0;
}

extern "C" int vk_cluster_rng_selftest(const uint32_t *key, int key_len, const int32_t *ns, int n_calls, int k,
                                       int32_t *out) {
    MT19937 rng;
    rng.init_by_array(key, key_len);
    std::vector<int32_t> pop, res;
    for (int i = 0; i < n_calls; ++i) {
        pop.resize((size_t)ns[i]);
        for (int j = 0; j < ns[i]; ++j) pop[(size_t)j] = j;
        const int kk = ns[i] < k ? ns[i] : k;
        rng.sample(pop, kk, res);
        for (int j = 0; j < k; ++j) out[(size_t)i * k + j] = j < kk ? res[(size_t)j] : -1;
    }
    return 0;
}

extern "C" int64_t vk_cluster_sizeof(int which) {
    return which == 0 ? (int64_t)sizeof(vk_cluster_config)
                      : (which == 1 ? (int64_t)sizeof(vk_cluster_result) : (int64_t)sizeof(vk_cluster_block));
}
