// Design aid (not product code, nothing links it): a timing model of the speculative sparse parse of
// flate_b200/csrc/lz77.cu on the host, to size ideas before spending GPU time on them.
//   - hash chains, findMatch and the lazy step as in the reference (simplified hash, no slide quirks: this is a
//     model of the WORK, not a parity oracle)
//   - a CTA = L lanes over a chunk of T positions (+ overlap W), one seed every G positions handed out last-first,
//     a lane stops at an arrival somebody claimed; cost of a search = ceil(candidates / 8) + 2 rounds
//   - event-driven: lanes advance in simulated time, so "who claims first" is as on the GPU
// Reports, per strategy, total work, makespan and lane utilisation = work / (L * makespan).
//   usage: sim_sparse FILE [level]
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint8_t* in;
static size_t n;
static int32_t* prev;
static int good = 8, lazy = 16, nice = 128, chain = 128;
static uint32_t hash4(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return (v * 0x9E3779B1u) >> 17; }

static long g_visits;
static int walk(size_t p, int min_len, int budget, int* dist, int* rounds) {
    int best = min_len, found = 0;
    size_t q = p;
    int maxl = n - p < 258 ? (int)(n - p) : 258;
    long v0 = g_visits;
    if (n - p >= 4) {
        while (budget-- > 0) {
            int32_t pr = prev[q];
            if (pr < 0) break;
            q = pr;
            if (p - q > 32768) break;
            g_visits++;
            int l = 0;
            while (l < maxl && in[q + l] == in[p + l]) l++;
            if (l > best && l >= 4) { best = l; *dist = (int)(p - q); found = 1; if (l >= nice) break; }
        }
    }
    *rounds += (int)((g_visits - v0 + 7) / 8) + 2;
    return found ? best : 0;
}
// lazy step from the clean arrival p: next clean arrival, cost in rounds
static size_t step(size_t p, int* rounds) {
    int d, l = walk(p, 0, chain, &d, rounds);
    if (!l) return p + 1;
    if (l >= lazy) return p + l;
    for (;;) {
        size_t q = p + 1;
        if (q >= n) return p + l;
        int d2, l2 = walk(q, l, l >= good ? chain >> 2 : chain, &d2, rounds);
        if (!l2) return p + l;
        p = q; l = l2;
        if (l >= lazy) return p + l;
    }
}

typedef struct { long t; int lane; } Ev;
static Ev* heap; static int hn;
static void hpush(Ev e) { int i = hn++; while (i && heap[(i - 1) / 2].t > e.t) { heap[i] = heap[(i - 1) / 2]; i = (i - 1) / 2; } heap[i] = e; }
static Ev hpop(void) { Ev top = heap[0], e = heap[--hn]; int i = 0; for (;;) { int c = 2 * i + 1; if (c >= hn) break; if (c + 1 < hn && heap[c + 1].t < heap[c].t) c++; if (heap[c].t >= e.t) break; heap[i] = heap[c]; i = c; } heap[i] = e; return top; }

// strategy 0: one seed list per chunk, L lanes, a lane takes the next seed when idle (as the kernel)
// strategy 1: as 0, plus: when no seeds are left, an idle lane splits the segment of the slowest running lane
//             (a new seed half-way between that lane's current arrival and the end of its segment)
static void run_chunks(int T, int W, int G, int L, int strategy, long* work_out, long* span_out, long* ideal_out) {
    uint8_t* claimed = calloc(n + 1024, 1);
    size_t* pos = malloc(sizeof(size_t) * L);       // current arrival of a lane
    size_t* seg_end = malloc(sizeof(size_t) * L);   // end of the segment it was seeded in
    char* busy = malloc(L);
    heap = malloc(sizeof(Ev) * (L + 8));
    long work = 0, span = 0, ideal = 0;
    for (size_t s = 0; s < n; s += T) {
        size_t span_end = s + T + W < n ? s + T + W : n;
        memset(claimed + s, 0, span_end - s);       // the next CTA evaluates its own copy of the overlap
        long nseeds = (long)((span_end - s + G - 1) / G), next_seed = 0, t_end = 0, chunk_work = 0;
        hn = 0;
        memset(busy, 0, L);
        for (int l = 0; l < L; l++) hpush((Ev){0, l});
        while (hn) {
            Ev e = hpop();
            int l = e.lane;
            size_t p;
            if (!busy[l]) {                          // needs a seed
                if (next_seed < nseeds) {
                    p = s + (size_t)(nseeds - 1 - next_seed) * G;
                    next_seed++;
                    seg_end[l] = p + G;
                } else if (strategy == 1) {
                    int best = -1; size_t gap = 8;   // split only if at least 8 positions are left
                    for (int k = 0; k < L; k++)
                        if (busy[k] && pos[k] < seg_end[k] && seg_end[k] - pos[k] > gap) { gap = seg_end[k] - pos[k]; best = k; }
                    if (best < 0) { if (e.t > t_end) t_end = e.t; continue; }
                    p = pos[best] + gap / 2;
                    seg_end[l] = seg_end[best];
                    seg_end[best] = p;
                } else { if (e.t > t_end) t_end = e.t; continue; }
                busy[l] = 1;
            } else p = pos[l];
            if (p >= span_end || claimed[p]) {       // left the span, or met somebody's trail
                busy[l] = 0;
                hpush((Ev){e.t + 1, l});
                continue;
            }
            claimed[p] = 1;
            int r = 0;
            pos[l] = step(p, &r);
            chunk_work += r;
            hpush((Ev){e.t + r, l});
        }
        work += chunk_work; span += t_end; ideal += (chunk_work + L - 1) / L;
    }
    *work_out = work; *span_out = span; *ideal_out = ideal;
    free(claimed); free(pos); free(seg_end); free(busy); free(heap);
}


// Rolling form: ONE persistent CTA over the whole input.  Ring of 65536 positions with load front lb (epochs of
// 4096); an arrival a may begin when a + 544 <= lb; an epoch may be loaded when min(active arrivals, next seed)
// >= lb + 4096 - 32768.  Seeds ascending.  A lane whose next arrival is not loaded parks it (list of `park_cap`
// entries; overflow = dropped, counted) and goes back to the seed counter; idle lanes resume parked arrivals first.
static void run_rolling(int G, int L, int park_cap) {
    uint8_t* claimed = calloc(n + 1024, 1);
    size_t* pos = malloc(sizeof(size_t) * L);
    char* busy = malloc(L);
    size_t* park = malloc(sizeof(size_t) * (park_cap + 1));
    int npark = 0;
    heap = malloc(sizeof(Ev) * (L + 8));
    hn = 0;
    long work = 0, t_end = 0, dropped = 0, parked_total = 0, waits = 0;
    size_t next_seed = 0, lb = 16384;
    const size_t load_end = (n + 544 + 4095) / 4096 * 4096;
    memset(busy, 0, L);
    for (int l = 0; l < L; l++) hpush((Ev){0, l});
    while (hn) {
        Ev e = hpop();
        int l = e.lane;
        // try to advance the load front as far as allowed (cheap in the model; on the GPU: one block-wide episode each)
        for (;;) {
            if (lb >= load_end) break;
            size_t pmin = next_seed < n ? next_seed : (size_t)-1;
            for (int k = 0; k < L; k++) if (busy[k] && pos[k] < pmin) pmin = pos[k];
            for (int k = 0; k < npark; k++) if (park[k] < pmin) pmin = park[k];
            // only load when somebody needs it: a parked arrival, or the seed front close to the loaded front
            int need = npark > 0 || next_seed + 2048 + 544 >= lb;
            if (!need || (pmin != (size_t)-1 && pmin + 28672 < lb)) break;
            lb += 4096;
        }
        const size_t avail = lb >= load_end ? n : lb - 544;
        size_t p;
        if (!busy[l]) {
            int got = 0;
            for (int k = 0; k < npark; k++)
                if (park[k] < avail) { p = park[k]; park[k] = park[--npark]; got = 1; break; }
            if (!got) {
                if (next_seed < n && next_seed < avail) { p = next_seed; next_seed += G; }
                else if (next_seed >= n && npark == 0) { if (e.t > t_end) t_end = e.t; continue; }  // nothing left for this lane
                else { waits++; hpush((Ev){e.t + 1, l}); continue; }                             // wait for a load
            }
            busy[l] = 1;
        } else p = pos[l];
        if (p >= n || claimed[p]) { busy[l] = 0; hpush((Ev){e.t + 1, l}); continue; }
        if (p >= avail) {  // next arrival not resident: park it and take other work
            if (npark < park_cap) { park[npark++] = p; parked_total++; } else dropped++;
            busy[l] = 0;
            hpush((Ev){e.t + 1, l});
            continue;
        }
        claimed[p] = 1;
        int r = 0;
        pos[l] = step(p, &r);
        work += r;
        hpush((Ev){e.t + r, l});
        if (e.t + r > t_end) t_end = e.t + r;
    }
    printf("rolling, G=%d, %d lanes, park list %d: work %ld rounds, makespan %ld, lane utilisation %.1f %%, parked %ld, dropped %ld, idle waits %ld\n",
           G, L, park_cap, work, t_end, 100.0 * work / ((double)L * t_end), parked_total, dropped, waits);
    free(claimed); free(pos); free(busy); free(park); free(heap);
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s FILE [level]\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END); n = ftell(f); fseek(f, 0, SEEK_SET);
    in = malloc(n + 8);
    if (fread(in, 1, n, f) != n) return 1;
    if (argc > 2) { int lv = atoi(argv[2]); if (lv == 9) { good = 32; lazy = 258; nice = 258; chain = 4096; } if (lv == 4) { good = 4; lazy = 4; nice = 16; chain = 16; } }
    prev = malloc(n * 4);
    int32_t* head = malloc(4 * 32768);
    memset(head, 0xff, 4 * 32768);
    for (size_t i = 0; i + 4 <= n; i++) { uint32_t h = hash4(in + i); prev[i] = head[h]; head[h] = (int32_t)i; }
    for (size_t i = n >= 4 ? n - 3 : 0; i < n; i++) prev[i] = -1;
    struct { int T, W, G, L, st; const char* name; } cfg[] = {
        {32768, 1024, 32, 1024, 0, "kernel as built (T=32768, G=32, 1024 lanes)"},
        {32768, 1024, 32, 1024, 1, "  + idle lanes split the slowest orbit's segment"},
        {32768, 1024, 16, 1024, 0, "G=16"},
        {32768, 1024, 32, 512, 0, "512 lanes"},
        {65536, 1024, 32, 1024, 0, "64 KiB per CTA"},
        {131072, 1024, 32, 1024, 0, "128 KiB per CTA"},
        {262144, 1024, 32, 1024, 0, "256 KiB per CTA"},
        {1 << 20, 1024, 32, 1024, 0, "1 MiB per CTA (what a rolling window approaches)"},
        {1 << 20, 1024, 64, 1024, 0, "1 MiB per CTA, G=64"},
        {1 << 20, 1024, 128, 1024, 0, "1 MiB per CTA, G=128"},
    };
    for (unsigned i = 0; i < sizeof cfg / sizeof cfg[0]; i++) {
        long w, sp, id;
        g_visits = 0;
        run_chunks(cfg[i].T, cfg[i].W, cfg[i].G, cfg[i].L, cfg[i].st, &w, &sp, &id);
        printf("%-58s work %9ld rounds, makespan %8ld, ideal %8ld, lane utilisation %.1f %%, candidates %ld\n", cfg[i].name, w, sp, id,
               100.0 * w / ((double)cfg[i].L * sp), g_visits);
    }
    run_rolling(32, 1024, 256);
    run_rolling(32, 768, 256);
    run_rolling(32, 512, 256);
    run_rolling(16, 1024, 256);
    run_rolling(16, 512, 256);
    run_rolling(8, 1024, 256);
    run_rolling(24, 1024, 256);
    return 0;
}
