/*
 * sharded_search.c — a plain-C caller of the drop-in library doing an 8-shard database search through the library alone
 * (LBAudioDetectiveDatabaseGroup*, include/LBAudioDetectiveDatabase.h): no torch, no NCCL, no transport of its own.  The reference's
 * callers are C programs that loop over archives with LBAudioDetectiveFingerprintCompareToFingerprint (FP.m:119-149,
 * LBAudioDetectiveTests.m:64-70); this is the multi-GPU form of that loop.
 *
 * The shards go to the visible CUDA devices round-robin (one device holds all eight on a single-GPU box).  Checks:
 *   - the group's top-k equals the top-k of ONE database holding all the clips, bit for bit (scores and global clip indices),
 *   - every query that is an excerpt of a database clip finds that clip with score 1,
 *   - clips added one at a time through fingerprint objects get consecutive global indices and are found.
 *
 * Build: cc -std=gnu11 -Iinclude tests/c/sharded_search.c -Llbaudiodetective_b200 -lLBAudioDetectiveCUDA -lm
 * Exit status 0 = all checks hold; 77 = no CUDA device (skipped).  argv[1] (optional) = clips in the database (default 20000).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "LBAudioDetective.h"
#include "LBAudioDetectiveDatabase.h"
#include "LBAudioDetectiveSupport.h"

#define SHARDS 8
#define L 200
#define W2 8            /* 2 * words per plane for L = 200 */
#define CLIP_SUBFPS 19
#define QUERY_SUBFPS 6
#define QUERIES 64
#define K 10

static unsigned long long rng_state = 88172645463325252ULL;
static unsigned next_u32(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (unsigned)(rng_state >> 16); }

/* a rank-sign code: every one of the 100 ranks carries exactly one of its two sign bits (what extraction produces) */
static void random_subfp(UInt32* w) {
    for (int i = 0; i < 4; i++) { UInt32 mask = i < 3 ? 0xffffffffu : 0xfu, r = next_u32() ^ (next_u32() << 16); w[i] = r & mask; w[4 + i] = ~r & mask; }
}

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { failures++; fprintf(stderr, "FAILED %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

int main(int argc, char** argv) {
    if (!LBAudioDetectiveSupportDeviceAvailable()) { fprintf(stderr, "no CUDA device: skipped\n"); return 77; }
    UInt32 clips = argc > 1 ? (UInt32)atoi(argv[1]) : 20000;
    int n_dev = LBAudioDetectiveSupportDeviceCount();
    CHECK(n_dev >= 1, "device count %d", n_dev);
    int devices[SHARDS];
    for (int i = 0; i < SHARDS; i++) devices[i] = i % n_dev;

    UInt32* words = malloc((size_t)clips * CLIP_SUBFPS * W2 * sizeof(UInt32));
    for (size_t i = 0; i < (size_t)clips * CLIP_SUBFPS; i++) random_subfp(words + i * W2);
    /* queries: excerpts of database clips at a random offset, every fourth one with a few sign flips */
    UInt32 qwords[QUERIES * QUERY_SUBFPS * W2], source[QUERIES];
    for (int q = 0; q < QUERIES; q++) {
        source[q] = next_u32() % clips; UInt32 off = next_u32() % (CLIP_SUBFPS - QUERY_SUBFPS + 1);
        memcpy(qwords + (size_t)q * QUERY_SUBFPS * W2, words + ((size_t)source[q] * CLIP_SUBFPS + off) * W2, QUERY_SUBFPS * W2 * sizeof(UInt32));
        if (q % 4 == 3) for (int f = 0; f < 5; f++) { UInt32* s = qwords + ((size_t)q * QUERY_SUBFPS + next_u32() % QUERY_SUBFPS) * W2; UInt32 bit = 1u << (next_u32() % 32); s[1] ^= bit; s[5] ^= bit; }
    }

    LBAudioDetectiveDatabaseGroupRef group = LBAudioDetectiveDatabaseGroupNew(L, devices, SHARDS);
    CHECK(group != NULL, "LBAudioDetectiveDatabaseGroupNew");
    if (!group) return 1;
    UInt64 first = 99;
    /* two appends, so that every shard holds two runs of global indices */
    UInt32 half = clips / 2;
    CHECK(LBAudioDetectiveDatabaseGroupAddPacked(group, words, half, NULL, CLIP_SUBFPS, &first) == noErr && first == 0, "first append");
    CHECK(LBAudioDetectiveDatabaseGroupAddPacked(group, words + (size_t)half * CLIP_SUBFPS * W2, clips - half, NULL, CLIP_SUBFPS, &first) == noErr && first == half, "second append");
    CHECK(LBAudioDetectiveDatabaseGroupGetNumberOfShards(group) == SHARDS && LBAudioDetectiveDatabaseGroupGetNumberOfClips(group) == clips, "group size");

    LBAudioDetectiveDatabaseRef whole = LBAudioDetectiveDatabaseNew(L);
    CHECK(whole != NULL && LBAudioDetectiveDatabaseAddPacked(whole, words, clips, NULL, CLIP_SUBFPS) == noErr, "single database");

    static Float32 g_sc[QUERIES * K], w_sc[QUERIES * K]; static UInt32 g_id[QUERIES * K], w_id[QUERIES * K];
    CHECK(LBAudioDetectiveDatabaseGroupSearchPacked(group, qwords, QUERIES, QUERY_SUBFPS, 0, K, g_sc, g_id) == noErr, "group search");
    CHECK(LBAudioDetectiveDatabaseSearchPacked(whole, qwords, QUERIES, QUERY_SUBFPS, 0, K, w_sc, w_id, NULL) == noErr, "single search");
    CHECK(memcmp(g_sc, w_sc, sizeof g_sc) == 0, "group scores differ from the single database's");
    CHECK(memcmp(g_id, w_id, sizeof g_id) == 0, "group clip indices differ from the single database's");
    int found = 0;
    for (int q = 0; q < QUERIES; q++) found += g_id[q * K] == source[q] && (q % 4 == 3 || g_sc[q * K] == 1.0f);
    CHECK(found == QUERIES, "%d of %d queries found their clip", found, QUERIES);
    printf("8-shard search over %u clips on %d device(s): top-%d of %d queries identical to one database, %d/%d excerpts found, %.3f ms on the device\n",
           (unsigned)clips, n_dev, K, QUERIES, found, QUERIES, LBAudioDetectiveDatabaseGroupGetLastSearchMilliseconds(group));

    /* one query (the server-style identify call) and a shortened range */
    CHECK(LBAudioDetectiveDatabaseGroupSearchPacked(group, qwords, 1, QUERY_SUBFPS, 77, K, g_sc, g_id) == noErr, "group search, one query");
    CHECK(LBAudioDetectiveDatabaseSearchPacked(whole, qwords, 1, QUERY_SUBFPS, 77, K, w_sc, w_id, NULL) == noErr, "single search, one query");
    CHECK(memcmp(g_sc, w_sc, K * sizeof(Float32)) == 0 && memcmp(g_id, w_id, K * sizeof(UInt32)) == 0, "one query, range 77: group differs");

    /* clips added one at a time as fingerprint objects */
    LBAudioDetectiveFingerprintRef fp = LBAudioDetectiveFingerprintNew(L);
    UInt32 sub[7 * W2];
    for (int i = 0; i < 7; i++) random_subfp(sub + i * W2);
    CHECK(LBAudioDetectiveFingerprintAddPackedSubfingerprints(fp, sub, 7) == noErr, "fingerprint");
    UInt64 id = 0;
    CHECK(LBAudioDetectiveDatabaseGroupAddFingerprint(group, fp, &id) == noErr && id == clips, "AddFingerprint index %llu", (unsigned long long)id);
    LBAudioDetectiveFingerprintRef query = LBAudioDetectiveFingerprintNew(L);
    LBAudioDetectiveFingerprintAddPackedSubfingerprints(query, sub + W2, QUERY_SUBFPS);
    CHECK(LBAudioDetectiveDatabaseGroupSearch(group, &query, 1, 0, 3, g_sc, g_id) == noErr && g_id[0] == clips && g_sc[0] == 1.0f, "fingerprint query: clip %u score %f", (unsigned)g_id[0], g_sc[0]);
    LBAudioDetectiveFingerprintDispose(fp); LBAudioDetectiveFingerprintDispose(query);

    LBAudioDetectiveDatabaseDispose(whole);
    LBAudioDetectiveDatabaseGroupDispose(group);
    free(words);
    if (failures) { fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
    printf("sharded search: all checks passed\n");
    return 0;
}
