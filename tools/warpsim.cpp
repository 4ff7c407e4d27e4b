// tools/warpsim.cpp -- DEVELOPMENT TOOL, NOT PRODUCT, NOT AN ORACLE.
//
// Lane-occupancy model of the trace kernel's warp loop (spb_wavefront.cu k_trace), run on the host
// with the host build of the device arithmetic (tests/hostsim): the bounce rays of sampled 8x4 pixel
// blocks of a frame are produced the way k_shade_hit_tiles produces them (shaded in item order, binned
// by direction), handed to simulated warps 32 slots at a time, and every iteration of the vote loop is
// counted with the number of lanes it runs on.  Lanes per instruction is a function of the algorithm
// and the ray order only, so scheduling policies can be compared here without a GPU; the model is
// checked against ncu's numbers for the production policy before anything is concluded from it
// (profiles/r2/README.md).
#include "../tests/hostsim/hostsim.cpp"

#include <algorithm>
#include <map>

namespace
{
struct SimRay
{
    f3 o, d;
    uint32_t rng;
    bool hole;
};

// instruction weights of the steps (SASS counts of the production kernel, rounded)
struct Weights
{
    double node = 145, leaf = 110, vote = 14, begin = 120, retire = 60, pop = 10, suspend = 40, resume = 40, perEntry = 4, enter = 250, exitStep = 90;
};

struct Policy
{
    uint32_t refillThreshold = 1; // 1: packet mode
    uint32_t speculate = 0;       // 1: one postponed leaf per lane
    uint32_t earlyLanes = 0;      // > 0: a packet with fewer walking lanes than this is suspended ...
    uint32_t earlyMinSteps = 0;   // ... once it has run this many iterations; survivors go to a second launch
    uint32_t secondThreshold = 1; // refill threshold of the second launch (continuations)
    uint32_t voteBias = 0;        // node step when nodeLanes + voteBias >= leafLanes
    uint32_t doubleNode = 0;      // > 0: after a node step that nodeLanes * 32 >= doubleNode * walking lanes chose, a second one without a vote
};

struct Tally
{
    double warpInst = 0, threadInst = 0;
    uint64_t nodeIters = 0, leafIters = 0, nodeLanes = 0, leafLanes = 0, rays = 0, hits = 0, nodeSteps = 0, leafSteps = 0;
    uint64_t suspended = 0, suspendedEntries = 0;
    uint64_t nodeLines = 0, leafLines = 0; // distinct 128-byte lines fetched per node / leaf iteration, summed
    uint64_t walkHist[33] = {};
    void add(double w, unsigned lanes)
    {
        warpInst += w;
        threadInst += w * lanes;
    }
};

const uint32_t NONE = 0xFFFFFFFFu;

struct Lane
{
    Trav st;
    TravCold cold;
    TravEntry stack[SPB_STACK_SIZE];
    v4f ray[2];
    bool have = false;
    uint32_t pending = NONE;
    uint32_t id = 0;
};

struct Continuation
{
    Trav st;
    TravCold cold;
    std::vector<TravEntry> stack;
    v4f ray[2];
    uint32_t pending;
    uint32_t id;
};

inline bool walking(const Lane &l) { return l.have && (trav_is_walking(l.st) || l.pending != NONE); }

// Moller-Trumbore on a postponed leaf (the triangle half of trav_leaf, without the pop)
inline void test_pending(const DScene &S, Lane &l)
{
    uint32_t index = l.pending & ~SPB_REF_LEAF;
    l.pending = NONE;
    const v4f *tp = S.tris + (size_t)index * 3;
    v4f a = ld4(tp + 0), b = ld4(tp + 1), cc = ld4(tp + 2);
    float t, u, v;
    if (ray_triangle_mt(l.st.o, l.st.d, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(cc.x, cc.y, cc.z), t, u, v))
        if (t > 0.0f && (t < l.st.lT || l.st.lT < 0.0f))
        {
            l.st.lT = t;
            l.st.lSlot = index;
            float c2 = t * SPB_CULL_SLACK;
            if (c2 < l.st.tcull) l.st.tcull = c2;
        }
}

// after a step of a speculating lane: park the leaf that came up and pop on
inline unsigned park_leaves(Lane &l)
{
    unsigned pops = 0;
    while (trav_is_walking(l.st) && !trav_is_node(l.st) && l.pending == NONE && l.st.blasBase >= 0)
    {
        l.pending = l.st.cur;
        trav_pop<true>(l.st, l.stack);
        pops++;
    }
    return pops;
}

struct Sim
{
    const DScene &S;
    Weights w;
    Policy p;
    Tally t;
    std::vector<Continuation> parked;
    std::vector<Hit> results; // by ray id
    bool single;
    uint64_t enterIters = 0, enterLanes = 0, exitIters = 0, exitLanes = 0;
    uint64_t spHist[97] = {}; // stack pointer after every node step
    Sim(const DScene &s, const Policy &pol, size_t rays) : S(s), p(pol), results(rays), single(s.objectCount == 1) {}
    bool finishedLane(const Lane &l) const { return l.have && l.pending == NONE && (l.st.cur == SPB_NODE_DONE || (single && l.st.cur == SPB_NODE_EXIT)); }

    void start(Lane &l, const SimRay &r, uint32_t id)
    {
        l.have = true;
        l.id = id;
        l.pending = NONE;
        l.ray[0].x = r.o.x; l.ray[0].y = r.o.y; l.ray[0].z = r.o.z; l.ray[0].w = 0;
        l.ray[1].x = r.d.x; l.ray[1].y = r.d.y; l.ray[1].z = r.d.z; l.ray[1].w = 0;
        if (single) trav_begin_single<true>(S, r.o, r.d, l.st, l.cold, nullptr);
        else trav_begin(S, r.o, r.d, l.st, l.cold);
        if (p.speculate) park_leaves(l);
    }
    void retire(Lane &l)
    {
        if (single && l.st.cur == SPB_NODE_EXIT) trav_leave(S, l.st, l.cold, l.ray);
        Hit h = trav_result(l.cold);
        if (l.cold.slow)
        {
            uint32_t stack[SPB_STACK_SIZE];
            float stackT[SPB_STACK_SIZE];
            h = intersect_scene<true, false>(S, mk3(l.ray[0].x, l.ray[0].y, l.ray[0].z), mk3(l.ray[1].x, l.ray[1].y, l.ray[1].z), stack, stackT, nullptr);
        }
        results[l.id] = h;
        t.rays++;
        if (h.t > 0.0f) t.hits++;
        l.have = false;
    }

    // one persistent warp over rays[begin, end) (or over continuations when `cont`)
    void run_warp(const std::vector<SimRay> &rays, size_t begin, size_t end, const std::vector<Continuation> *cont, uint32_t threshold)
    {
        std::vector<Lane> lanes(32);
        size_t next = begin;
        bool exhausted = begin >= end;
        for (;;)
        {
            unsigned fin = 0;
            for (auto &l : lanes)
                if (finishedLane(l)) { retire(l); fin++; }
            if (fin) t.add(w.retire, fin);
            if (!exhausted)
            {
                unsigned filled = 0, entries = 0;
                for (auto &l : lanes)
                {
                    if (l.have) continue;
                    if (next >= end) break;
                    size_t idx = next++;
                    if (cont)
                    {
                        const Continuation &c = (*cont)[idx];
                        l.st = c.st; l.cold = c.cold; l.ray[0] = c.ray[0]; l.ray[1] = c.ray[1];
                        l.pending = c.pending; l.id = c.id; l.have = true;
                        std::copy(c.stack.begin(), c.stack.end(), l.stack);
                        entries += (unsigned)c.stack.size();
                        filled++;
                    }
                    else if (!rays[idx].hole)
                    {
                        start(l, rays[idx], (uint32_t)idx);
                        filled++;
                    }
                }
                if (next >= end) exhausted = true;
                if (filled) t.add(cont ? w.resume : w.begin, filled);
                if (cont && entries) t.add(w.perEntry * entries / std::max(1u, filled), filled);
            }
            if (!single)
            {
                // lanes whose walk inside an object has ended leave it together (outer loop of k_trace)
                unsigned ex = 0;
                for (auto &l : lanes)
                    if (l.have && l.st.cur == SPB_NODE_EXIT && l.pending == NONE) { trav_exit<true>(S, l.st, l.cold, l.ray, l.stack); ex++; }
                if (ex) { t.add(w.exitStep, ex); exitIters++; exitLanes += ex; }
            }
            unsigned nwalk = 0, nhave = 0;
            for (auto &l : lanes) { nwalk += walking(l); nhave += l.have; }
            if (!nwalk)
            {
                if (!nhave && exhausted) break;
                continue;
            }
            uint32_t iters = 0;
            do
            {
                unsigned nodeLanes = 0, leafLanes = 0;
                for (auto &l : lanes)
                {
                    if (!walking(l)) continue;
                    bool node = trav_is_walking(l.st) && trav_is_node(l.st);
                    if (node) nodeLanes++;
                    if (p.speculate ? (l.pending != NONE || (trav_is_walking(l.st) && !node)) : !node) leafLanes++;
                }
                unsigned blocked = p.speculate ? nwalk - nodeLanes : leafLanes;
                t.walkHist[nwalk]++;
                t.add(w.vote, 32);
                // (voteBias 1000: the step that leaves the fewest lane-slots idle -- cost x (32 - lanes) -- instead of the one with more lanes;
                //  1001: the same with equal costs, i.e. the plain majority but without the preference for nodes on a tie)
                const bool chooseNode = p.voteBias == 1000 ? (nodeLanes && (!blocked || w.node * (32 - nodeLanes) <= w.leaf * (32 - blocked)))
                                        : p.voteBias == 1001 ? (nodeLanes && nodeLanes > blocked)
                                                             : (nodeLanes + p.voteBias >= blocked && nodeLanes);
                if (chooseNode)
                {
                    unsigned pops = 0;
                    {
                        uint32_t seen[32]; unsigned ns = 0;
                        for (auto &l : lanes)
                            if (walking(l) && trav_is_walking(l.st) && trav_is_node(l.st))
                            {
                                bool dup = false;
                                for (unsigned k = 0; k < ns; ++k) dup |= seen[k] == l.st.cur;
                                if (!dup) seen[ns++] = l.st.cur;
                            }
                        t.nodeLines += ns;
                    }
                    for (auto &l : lanes)
                        if (walking(l) && trav_is_walking(l.st) && trav_is_node(l.st))
                        {
                            trav_node<true>(S, l.st, l.stack, nullptr);
                            spHist[l.st.sp < 96 ? l.st.sp : 96]++;
                            if (p.speculate) pops += park_leaves(l);
                        }
                    t.add(w.node, nodeLanes);
                    if (pops) t.add(w.pop, std::min(pops, nodeLanes));
                    t.nodeIters++; t.nodeLanes += nodeLanes; t.nodeSteps += nodeLanes;
                    if (p.doubleNode && nodeLanes * 32 >= p.doubleNode * nwalk)
                    {
                        unsigned again = 0;
                        for (auto &l : lanes)
                            if (walking(l) && trav_is_walking(l.st) && trav_is_node(l.st))
                            {
                                trav_node<true>(S, l.st, l.stack, nullptr);
                                again++;
                            }
                        if (again) { t.add(w.node + 4, again); t.nodeIters++; t.nodeLanes += again; t.nodeSteps += again; }
                        else t.add(4, 32);
                    }
                }
                else
                {
                    unsigned pops = 0, entering = 0;
                    {
                        // triangles are 48 bytes: count distinct 128-byte lines touched
                        uint32_t seen[64]; unsigned ns = 0;
                        for (auto &l : lanes)
                        {
                            if (!walking(l)) continue;
                            uint32_t ref = p.speculate ? l.pending : (trav_is_node(l.st) ? NONE : l.st.cur);
                            if (ref == NONE) continue;
                            uint32_t b0 = (ref & ~SPB_REF_LEAF) * 48u, lines[2] = {b0 / 128u, (b0 + 47u) / 128u};
                            for (int q = 0; q < 2; ++q)
                            {
                                bool dup = false;
                                for (unsigned k = 0; k < ns; ++k) dup |= seen[k] == lines[q];
                                if (!dup) seen[ns++] = lines[q];
                            }
                        }
                        t.leafLines += ns;
                    }
                    for (auto &l : lanes)
                    {
                        if (!walking(l)) continue;
                        if (p.speculate)
                        {
                            if (l.pending != NONE) { test_pending(S, l); pops += park_leaves(l); }
                            else if (trav_is_walking(l.st) && !trav_is_node(l.st))
                            {
                                if (l.st.blasBase < 0) entering++;
                                trav_leaf<true>(S, l.st, l.cold, l.ray, l.stack, nullptr);
                                pops += park_leaves(l);
                            }
                        }
                        else if (trav_is_walking(l.st) && !trav_is_node(l.st))
                        {
                            if (l.st.blasBase < 0) entering++;
                            trav_leaf<true>(S, l.st, l.cold, l.ray, l.stack, nullptr);
                        }
                    }
                    // the two branches of trav_leaf run one after the other when a warp has both kinds of lane
                    if (leafLanes > entering) t.add(w.leaf, leafLanes - entering);
                    if (entering) { t.add(w.enter, entering); enterIters++; enterLanes += entering; }
                    if (pops) t.add(w.pop, std::min(pops, leafLanes));
                    t.leafIters++; t.leafLanes += leafLanes; t.leafSteps += leafLanes;
                }
                nwalk = 0;
                for (auto &l : lanes) nwalk += walking(l);
                iters++;
                if (!cont && p.earlyLanes && nwalk && nwalk < p.earlyLanes && iters >= p.earlyMinSteps)
                {
                    // suspend the survivors: state + live stack entries to the continuation buffer
                    unsigned entries = 0;
                    for (auto &l : lanes)
                        if (walking(l))
                        {
                            Continuation c;
                            c.st = l.st; c.cold = l.cold; c.ray[0] = l.ray[0]; c.ray[1] = l.ray[1]; c.pending = l.pending; c.id = l.id;
                            c.stack.assign(l.stack, l.stack + l.st.sp);
                            entries += (unsigned)l.st.sp;
                            parked.push_back(c);
                            l.have = false;
                            t.suspended++;
                        }
                    t.suspendedEntries += entries;
                    t.add(w.suspend, nwalk);
                    t.add(w.perEntry * entries / nwalk, nwalk);
                    nwalk = 0;
                }
            } while (nwalk && (nwalk >= threshold || exhausted));
        }
    }
};

// direction bin of k_shade_hit_tiles (spb_wavefront.cu direction_bin, 256 bins)
unsigned g_binRes = 16;
unsigned direction_bin(f3 d)
{
    float n = fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
    float inv = n > 0.0f ? 1.0f / n : 0.0f;
    float u = d.x * inv, v = d.y * inv;
    if (d.z < 0.0f)
    {
        float fu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f);
        float fv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f);
        u = fu; v = fv;
    }
    const int R = (int)g_binRes;
    int iu = (int)((u * 0.5f + 0.5f) * (float)R), iv = (int)((v * 0.5f + 0.5f) * (float)R);
    unsigned qu = (unsigned)(iu < 0 ? 0 : (iu > R - 1 ? R - 1 : iu)), qv = (unsigned)(iv < 0 ? 0 : (iv > R - 1 ? R - 1 : iv));
    // Morton interleave of two coordinates of up to 8 bits
    auto spread = [](unsigned x) { x = (x | (x << 4)) & 0x0F0Fu; x = (x | (x << 2)) & 0x3333u; x = (x | (x << 1)) & 0x5555u; return x; };
    return spread(qu) | (spread(qv) << 1);
}

// the next ray of a hit (shade_hit_one without the BSDF terms)
bool bounce_ray(const DScene &S, const SimRay &r, Hit h, SimRay &out)
{
    if (!(h.t > 0.0f)) return false;
    hit_barycentrics(S, r.o, r.d, h);
    Surface sf = resolve_hit(S, h);
    f3 P = add3(r.o, mul3(r.d, h.t));
    uint32_t rng = r.rng;
    f3 L = random_hemisphere<0>(sf.normal, rng);
    out.o = add3(P, mul3(sf.normal, 0.0001f));
    out.d = L;
    out.rng = rng;
    out.hole = false;
    return true;
}
} // namespace

// out: per launch (bounce 1, bounce 2, ...) 16 doubles:
//  [0] rays [1] hits [2] warp instructions [3] thread instructions [4] node iterations [5] node lanes
//  [6] leaf iterations [7] leaf lanes [8] suspended rays [9] suspended stack entries [10] second-launch warp inst
//  [11] second-launch thread inst [12] node steps [13] leaf steps
// hist: 33 doubles of the first launch: vote iterations by walking-lane count
extern "C" void warpsim_run(ora_Scene *s, uint32_t blockStep, uint32_t spp, uint32_t bounces, uint32_t frame,
                            const uint32_t *policy, const double *weights, double *out, double *hist, uint32_t sortLater)
{
    refresh(s);
    const DScene &S = s->d;
    const DCamera &c = s->dc;
    Policy pol;
    pol.refillThreshold = policy[0]; pol.speculate = policy[1]; pol.earlyLanes = policy[2]; pol.earlyMinSteps = policy[3];
    pol.secondThreshold = policy[4]; pol.voteBias = policy[5];
    const uint32_t laterThreshold = policy[6];
    g_binRes = policy[7] ? policy[7] : 16;
    const uint32_t pixelMajor = policy[8];
    pol.doubleNode = policy[9];
    Weights w;
    if (weights) { w.node = weights[0]; w.leaf = weights[1]; w.vote = weights[2]; w.begin = weights[3]; w.retire = weights[4]; w.pop = weights[5]; w.suspend = weights[6]; w.resume = weights[7]; w.perEntry = weights[8]; w.enter = weights[9]; w.exitStep = weights[10]; }

    // primary pass over the sampled tiles -> bounce-1 queue.  A tile is SPB_SORT_TILE = 2048 consecutive items:
    // 2048 / (32 * spp) consecutive blocks of the covered-block list (row-major), all their pixels and samples.
    std::vector<SimRay> queue;
    const uint32_t blocksX = (c.width + 7) / 8, blocksY = (c.height + 3) / 4;
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    const uint32_t tileBlocks = std::max(1u, 2048u / (32u * spp));
    std::vector<std::pair<uint32_t, uint32_t>> group;
    uint32_t counter = 0;
    auto flush_tile = [&]() {
        if (group.empty()) return;
        const bool take = (counter++ % blockStep) == 0;
        std::vector<std::pair<uint32_t, uint32_t>> blocks;
        blocks.swap(group);
        if (!take) return;
        std::vector<SimRay> tile;
        std::vector<unsigned> bins;
        for (auto &bb : blocks)
            for (uint32_t l = 0; l < 32; ++l)
                for (uint32_t sample = 0; sample < spp; ++sample)
                {
                    uint32_t x = bb.first * 8 + (l & 7u), y = bb.second * 4 + (l >> 3);
                    if (x >= c.width || y >= c.height) continue;
                    SimRay r;
                    r.rng = stream_seed(x + y * c.width, sample, frame);
                    primary_ray(c, x, y, r.rng, r.o, r.d);
                    r.hole = false;
                    Hit h = intersect_scene_stepped<true>(S, r.o, r.d, stack, stackT, nullptr);
                    SimRay nr;
                    if (bounce_ray(S, r, h, nr))
                    {
                        tile.push_back(nr);
                        // pixelMajor: 0 = direction bin only; k = pixel group (l / k) above the direction bin
                        bins.push_back(direction_bin(nr.d) + (pixelMajor ? (l / pixelMajor) * 65536u : 0u));
                    }
                }
        // bin by direction, item order inside a bin; the tile's unused slots are holes
        std::vector<size_t> order(tile.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        if (sortLater & 2u)
        {
            // (sensitivity check) random order inside a bin
            uint32_t rs = 0x9E3779B9u ^ (blocks[0].first * 7919u + blocks[0].second);
            for (size_t i = order.size(); i > 1; --i) { xorshift32(rs); std::swap(order[i - 1], order[rs % i]); }
        }
        if (!(sortLater & 4u)) std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return bins[a] < bins[b]; });
        size_t tileSlots = (size_t)32 * spp * tileBlocks;
        for (size_t i = 0; i < tileSlots; ++i)
        {
            SimRay r;
            r.hole = true;
            r.o = r.d = mk3(0, 0, 0);
            r.rng = 0;
            if (i < order.size()) r = tile[order[i]];
            queue.push_back(r);
        }
    };
    for (uint32_t by = 0; by < blocksY; ++by)
        for (uint32_t bx = 0; bx < blocksX; ++bx)
        {
            // blocks with geometry: sample 0 of every pixel (the device lists every block a padded triangle
            // box touches; blocks without a single hit contribute no bounce rays either way)
            bool any = false;
            for (uint32_t l = 0; l < 32 && !any; ++l)
            {
                uint32_t x = bx * 8 + (l & 7u), y = by * 4 + (l >> 3);
                if (x >= c.width || y >= c.height) continue;
                uint32_t rng = stream_seed(x + y * c.width, 0, frame);
                f3 o, d;
                primary_ray(c, x, y, rng, o, d);
                Hit h = intersect_scene_stepped<true>(S, o, d, stack, stackT, nullptr);
                any = h.t > 0.0f;
            }
            if (!any) continue;
            group.push_back({bx, by});
            if (group.size() == tileBlocks) flush_tile();
        }
    flush_tile();

    for (uint32_t bounce = 1; bounce < bounces; ++bounce)
    {
        Policy pl = pol;
        if (bounce > 1 && !(sortLater & 1u)) { pl.refillThreshold = laterThreshold; }
        Sim sim(S, pl, queue.size());
        sim.w = w;
        // warps of the launch: packet mode hands a warp 32 consecutive slots at a time, so one simulated warp per
        // 64-slot chunk pair is equivalent to the device's chunked hand-out; in threshold mode a warp streams
        // through a longer range (the device interleaves chunks of many warps; coherence does not enter this model)
        const size_t span = pl.refillThreshold <= 1 ? 2048 : 8192;
        for (size_t b = 0; b < queue.size(); b += span) sim.run_warp(queue, b, std::min(queue.size(), b + span), nullptr, pl.refillThreshold);
        Tally first = sim.t;
        double w2 = 0, t2 = 0;
        if (!sim.parked.empty())
        {
            std::vector<Continuation> cont;
            cont.swap(sim.parked);
            Tally before = sim.t;
            const size_t span2 = pl.secondThreshold <= 1 ? 2048 : 8192;
            for (size_t b = 0; b < cont.size(); b += span2) sim.run_warp(queue, b, std::min(cont.size(), b + span2), &cont, pl.secondThreshold);
            w2 = sim.t.warpInst - before.warpInst;
            t2 = sim.t.threadInst - before.threadInst;
        }
        double *o = out + (size_t)(bounce - 1) * 24;
        o[0] = (double)sim.t.rays; o[1] = (double)sim.t.hits; o[2] = sim.t.warpInst; o[3] = sim.t.threadInst;
        o[4] = (double)sim.t.nodeIters; o[5] = (double)sim.t.nodeLanes; o[6] = (double)sim.t.leafIters; o[7] = (double)sim.t.leafLanes;
        o[8] = (double)sim.t.suspended; o[9] = (double)sim.t.suspendedEntries; o[10] = w2; o[11] = t2;
        o[12] = (double)sim.t.nodeSteps; o[13] = (double)sim.t.leafSteps; o[14] = (double)sim.t.nodeLines; o[15] = (double)sim.t.leafLines;
        o[16] = (double)sim.enterIters; o[17] = (double)sim.enterLanes; o[18] = (double)sim.exitIters; o[19] = (double)sim.exitLanes;
        if (bounce == 1 && hist)
            for (int i = 0; i < 33; ++i) hist[i] = (double)first.walkHist[i];
        if (getenv("WARPSIM_SP"))
        {
            double tot = 0, acc = 0;
            for (int i = 0; i < 97; ++i) tot += (double)sim.spHist[i];
            fprintf(stderr, "bounce %u: stack pointer after a node step, cumulative:", bounce);
            for (int i = 0; i < 40; ++i) { acc += (double)sim.spHist[i]; fprintf(stderr, " %d:%.4f", i, acc / tot); }
            fprintf(stderr, "\n");
        }
        // next queue: the hits in retire order ~ slot order (k_shade_hit: slot for slot)
        std::vector<SimRay> nextq;
        for (size_t i = 0; i < queue.size(); ++i)
        {
            if (queue[i].hole) continue;
            SimRay nr;
            if (bounce_ray(S, queue[i], sim.results[i], nr)) nextq.push_back(nr);
        }
        queue.swap(nextq);
        if (queue.empty()) break;
    }
}
