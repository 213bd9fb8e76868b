"""CPU: the work decomposition and the producer's ring protocol of the chain kernel (csrc/gemm_chain.cu), replayed on the
host from the same geometry header (csrc/chain_geo.h, compiled with gcc).  For several layer stacks, cluster counts and
cluster sizes: every (layer, tile, k-block) is owned by exactly one CTA, every CTA of a cluster walks the same item list
(the two cluster barriers per item stay matched), every CTA arrives once per layer, and the producer's weight-tile
prefetch never re-arms a ring slot whose occupant is still waiting for its activation tile (that would be a deadlock:
the MMA that frees the slot could never run)."""
import os
import subprocess

import pytest

from conftest import ROOT

SRC = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "chain_geo.h"
#define STAGES 7
#define MAXL 16
static int N[MAXL], K[MAXL], L;
static void normalize(int cid, int* l, int* t) { while (*l < L && *t >= chain_tiles(N[*l])) { ++*l; *t = cid; } }
int main(int argc, char** argv) {
    int C = atoi(argv[1]), S = atoi(argv[2]);
    L = (argc - 3) / 2;
    for (int i = 0; i < L; ++i) { N[i] = atoi(argv[3 + 2 * i]); K[i] = atoi(argv[4 + 2 * i]); }
    static int owned[MAXL][80][80];               /* [layer][tile][k-block] -> number of owners */
    memset(owned, 0, sizeof(owned));
    int arrivals[MAXL] = {0};
    long items_of_cluster[80] = {0};
    for (int cid = 0; cid < C; ++cid)
        for (int rank = 0; rank < S; ++rank) {
            int l = 0, t = cid, arrived = 0, items = 0;
            /* producer ring: a slot may be re-armed only when the k-block in it has had its A tile issued */
            int slot_item[STAGES], slot_has_a[STAGES], ring_b = 0, ring_a = 0, pre_done = 0;
            for (int s = 0; s < STAGES; ++s) { slot_item[s] = -1; slot_has_a[s] = 1; }
            normalize(cid, &l, &t);
            if (l < L) {
                struct Geo g0 = chain_layer_geo(N[l], K[l], S, rank);
                pre_done = g0.num_kb < STAGES ? g0.num_kb : STAGES;
                for (int kb = 0; kb < pre_done; ++kb) {
                    if (!slot_has_a[ring_b]) { printf("DEADLOCK prologue\n"); return 1; }
                    slot_item[ring_b] = 0; slot_has_a[ring_b] = 0; ring_b = (ring_b + 1) % STAGES;
                }
            }
            while (l < L) {
                struct Geo g = chain_layer_geo(N[l], K[l], S, rank);
                int nl = l, nt = t + C;
                normalize(cid, &nl, &nt);
                while (arrived < l) { ++arrivals[arrived]; ++arrived; }
                if (g.nsplit > S || g.nsplit < 1) { printf("BAD nsplit\n"); return 1; }
                if ((rank >= g.nsplit) != (g.num_kb == 0)) { printf("BAD empty rank\n"); return 1; }
                for (int kb = 0; kb < g.num_kb; ++kb) {
                    ++owned[l][t][g.kb_begin + kb];
                    if (kb >= pre_done) {
                        if (!slot_has_a[ring_b]) { printf("DEADLOCK in-item\n"); return 1; }
                        slot_item[ring_b] = items; slot_has_a[ring_b] = 0; ring_b = (ring_b + 1) % STAGES;
                    }
                    if (slot_item[ring_a] != items || slot_has_a[ring_a]) { printf("RING ORDER\n"); return 1; }
                    slot_has_a[ring_a] = 1; ring_a = (ring_a + 1) % STAGES;
                }
                pre_done = 0;
                if (nl < L) {
                    struct Geo g2 = chain_layer_geo(N[nl], K[nl], S, rank);
                    pre_done = g2.num_kb < STAGES ? g2.num_kb : STAGES;
                    for (int kb = 0; kb < pre_done; ++kb) {
                        if (!slot_has_a[ring_b]) { printf("DEADLOCK prefetch\n"); return 1; }
                        slot_item[ring_b] = items + 1; slot_has_a[ring_b] = 0; ring_b = (ring_b + 1) % STAGES;
                    }
                }
                ++items;
                l = nl; t = nt;
            }
            while (arrived < L - 1) { ++arrivals[arrived]; ++arrived; }
            if (rank == 0) items_of_cluster[cid] = items;
            else if (items_of_cluster[cid] != items) { printf("CLUSTER ITEM COUNT MISMATCH\n"); return 1; }
        }
    for (int l = 0; l < L; ++l) {
        int tiles = chain_tiles(N[l]), kbs = (K[l] + 63) / 64;
        for (int t = 0; t < tiles; ++t)
            for (int kb = 0; kb < kbs; ++kb)
                if (owned[l][t][kb] != 1) { printf("COVERAGE layer %d tile %d kb %d owners %d\n", l, t, kb, owned[l][t][kb]); return 1; }
        if (l < L - 1 && arrivals[l] != C * S) { printf("ARRIVALS layer %d: %d != %d\n", l, arrivals[l], C * S); return 1; }
    }
    printf("OK\n");
    return 0;
}
'''

# (N, K) per layer
STACKS = {
    "embedding": [(1536, 1537)] * 10,
    "bottleneck": [(1536, 1537), (1067, 1537), (598, 1073), (128, 601), (597, 129), (1066, 601), (1535, 1073), (1536, 1537)],
    "small": [(192, 193)] * 8,
    "dgrad": [(1536, 1536)] * 9,
    "wide_then_narrow": [(4096, 1537), (64, 4097), (2048, 65)],
}


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    d = tmp_path_factory.mktemp("chain")
    (d / "sim.c").write_text(SRC)
    exe = d / "sim"
    r = subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "mui-deepautoencoder_b200", "csrc"), str(d / "sim.c"), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(exe)


@pytest.mark.parametrize("stack", sorted(STACKS))
@pytest.mark.parametrize("C,S", [(24, 6), (3, 8), (1, 1), (7, 5), (18, 8), (64, 2)])
def test_every_k_block_has_one_owner_and_the_protocol_cannot_deadlock(sim, stack, C, S):
    args = [str(v) for nk in STACKS[stack] for v in nk]
    r = subprocess.run([sim, str(C), str(S)] + args, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK", (stack, C, S, r.stdout)
