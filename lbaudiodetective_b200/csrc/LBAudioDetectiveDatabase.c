/*
 * LBAudioDetectiveDatabase.c — host side of the batched matcher (see include/LBAudioDetectiveDatabase.h).
 * Marshals packed fingerprints into the CUDA layer (lbad_search.cu); no match arithmetic happens here.
 */
#include "lbad_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct LBAudioDetectiveDatabase {
    UInt32 subfingerprintLength;
    UInt32 W;
    lbadcu_db* db;
};

LBAudioDetectiveDatabaseRef LBAudioDetectiveDatabaseNew(UInt32 inSubfingerprintLength) {
    UInt32 W = lbad_words_per_plane(inSubfingerprintLength);
    if (!W) return NULL;
    LBAudioDetectiveDatabaseRef d = calloc(1, sizeof *d);
    if (!d) return NULL;
    d->subfingerprintLength = inSubfingerprintLength; d->W = W;
    if (lbadcu_db_create(W, (inSubfingerprintLength + 1) / 2, &d->db) != LBAD_OK) { free(d); return NULL; }
    return d;
}

OSStatus LBAudioDetectiveDatabaseDispose(LBAudioDetectiveDatabaseRef d) {
    if (!d) return kLBAudioDetectiveArgumentInvalid;
    lbadcu_db_destroy(d->db); free(d);
    return noErr;
}

UInt32 LBAudioDetectiveDatabaseGetNumberOfClips(LBAudioDetectiveDatabaseRef d) { return d ? lbadcu_db_clips(d->db) : 0; }
UInt64 LBAudioDetectiveDatabaseGetNumberOfSubfingerprints(LBAudioDetectiveDatabaseRef d) { return d ? lbadcu_db_subfps(d->db) : 0; }

OSStatus LBAudioDetectiveDatabaseSetClipIndexBase(LBAudioDetectiveDatabaseRef d, UInt32 inBase) {
    if (!d) return kLBAudioDetectiveArgumentInvalid;
    lbadcu_db_set_base(d->db, inBase);
    return noErr;
}

OSStatus LBAudioDetectiveDatabaseAddFingerprint(LBAudioDetectiveDatabaseRef d, LBAudioDetectiveFingerprintRef fp, UInt32* outClipIndex) {
    if (!d || !fp || fp->subfingerprintLength != d->subfingerprintLength) return kLBAudioDetectiveArgumentInvalid;
    UInt32 idx = lbadcu_db_clips(d->db), n = fp->subfingerprintCount, dummy[16] = {0};
    OSStatus e = lbad_status(lbadcu_db_append(d->db, n ? fp->words : dummy, 0, 1, &n, 0, NULL, -1));
    if (e == noErr && outClipIndex) *outClipIndex = idx;
    return e;
}

OSStatus LBAudioDetectiveDatabaseAddPacked(LBAudioDetectiveDatabaseRef d, const UInt32* inWords, UInt32 nClips, const UInt32* inCounts, UInt32 uniform) {
    if (!d || !inWords) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_db_append(d->db, inWords, 0, nClips, inCounts, uniform, NULL, -1));
}

OSStatus LBAudioDetectiveDatabaseAddPackedDevice(LBAudioDetectiveDatabaseRef d, const UInt32* inDeviceWords, UInt32 nClips, UInt32 uniform, void* inProducerStream) {
    if (!d || !inDeviceWords) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_db_append(d->db, inDeviceWords, 1, nClips, NULL, uniform, inProducerStream, -1));
}

static UInt32 pairs_for(LBAudioDetectiveDatabaseRef d, UInt32 inRange) {
    if (inRange == 0) inRange = d->subfingerprintLength;                         /* LBAudioDetective.m:443-445 */
    return lbad_pairs_for_range(inRange, d->subfingerprintLength);
}

OSStatus LBAudioDetectiveDatabaseSearchPacked(LBAudioDetectiveDatabaseRef d, const UInt32* inQueryWords, UInt32 nQ, UInt32 qCount, UInt32 inRange, UInt32 inK,
                                              Float32* outScores, UInt32* outClipIndices, Float32* outAllScores) {
    if (!d || (!inQueryWords && qCount) || !outScores || !outClipIndices) return kLBAudioDetectiveArgumentInvalid;
    UInt32 dummy[16] = {0};
    return lbad_status(lbadcu_db_search_host(d->db, qCount ? inQueryWords : dummy, nQ, qCount, pairs_for(d, inRange), inK, outScores, outClipIndices, outAllScores));
}

OSStatus LBAudioDetectiveDatabaseSearch(LBAudioDetectiveDatabaseRef d, const LBAudioDetectiveFingerprintRef* inQueries, UInt32 nQ, UInt32 inRange, UInt32 inK,
                                        Float32* outScores, UInt32* outClipIndices) {
    if (!d || !inQueries || nQ == 0) return kLBAudioDetectiveArgumentInvalid;
    UInt32 qCount = inQueries[0]->subfingerprintCount;
    for (UInt32 i = 0; i < nQ; i++)
        if (inQueries[i]->subfingerprintCount != qCount || inQueries[i]->subfingerprintLength != d->subfingerprintLength) return kLBAudioDetectiveArgumentInvalid;
    size_t per = (size_t)qCount * 2 * d->W;
    UInt32* words = malloc(((per * nQ) > 0 ? per * nQ : 1) * sizeof(UInt32));
    if (!words) return kLBAudioDetectiveArgumentInvalid;
    for (UInt32 i = 0; i < nQ; i++) memcpy(words + per * i, inQueries[i]->words, per * sizeof(UInt32));
    OSStatus e = LBAudioDetectiveDatabaseSearchPacked(d, words, nQ, qCount, inRange, inK, outScores, outClipIndices, NULL);
    free(words);
    return e;
}

OSStatus LBAudioDetectiveDatabaseSearchDevice(LBAudioDetectiveDatabaseRef d, const UInt32* dQ, UInt32 nQ, UInt32 qCount, UInt32 inRange, UInt32 inK,
                                              Float32* dScores, UInt32* dIdx, void* stream) {
    if (!d || !dQ || !dScores || !dIdx) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_db_search_device(d->db, dQ, nQ, qCount, pairs_for(d, inRange), inK, dScores, dIdx, NULL, stream));
}

OSStatus LBAudioDetectiveDatabaseMergeTopK(const Float32* inScores, const UInt32* inClipIndices, UInt32 nLists, UInt32 nQ, UInt32 inK, Float32* outScores, UInt32* outClipIndices) {
    return lbad_status(lbadcu_merge_topk_host(inScores, inClipIndices, nLists, nQ, inK, outScores, outClipIndices));
}

OSStatus LBAudioDetectiveDatabaseMergeTopKDevice(const Float32* dScores, const UInt32* dIdx, UInt32 nLists, UInt32 nQ, UInt32 inK, Float32* dOutScores, UInt32* dOutIdx, void* stream) {
    return lbad_status(lbadcu_merge_topk_device(dScores, dIdx, nLists, 0, nQ, inK, dOutScores, dOutIdx, stream));
}

OSStatus LBAudioDetectiveDatabaseMergeTopKDeviceStrided(const Float32* dScores, const UInt32* dIdx, UInt32 nLists, UInt64 inListStride, UInt32 nQ, UInt32 inK, Float32* dOutScores, UInt32* dOutIdx, void* stream) {
    return lbad_status(lbadcu_merge_topk_device(dScores, dIdx, nLists, inListStride, nQ, inK, dOutScores, dOutIdx, stream));
}

UInt64 LBAudioDetectiveDatabaseComparesPerQuery(LBAudioDetectiveDatabaseRef d, UInt32 qCount) { return d ? lbadcu_db_compares_per_query(d->db, qCount) : 0; }
UInt64 LBAudioDetectiveDatabaseGetKernelLaunchCount(LBAudioDetectiveDatabaseRef d) { return d ? lbadcu_db_launches(d->db) : 0; }
UInt32 LBAudioDetectiveDatabaseGetKernelTiming(LBAudioDetectiveDatabaseRef d, Boolean inEnable, Boolean inReset, Float64* outTotalMilliseconds) {
    if (outTotalMilliseconds) *outTotalMilliseconds = 0.0;
    return d ? lbadcu_db_timing(d->db, inEnable, inReset, outTotalMilliseconds) : 0;
}

/* ---- group: one process, several GPUs (include/LBAudioDetectiveDatabase.h) ---- */

struct LBAudioDetectiveDatabaseGroup {
    UInt32 subfingerprintLength;
    UInt32 W;
    lbadcu_group* group;
};

LBAudioDetectiveDatabaseGroupRef LBAudioDetectiveDatabaseGroupNew(UInt32 inSubfingerprintLength, const int* inDevices, UInt32 inNumberOfShards) {
    UInt32 W = lbad_words_per_plane(inSubfingerprintLength);
    if (!W || !inDevices || inNumberOfShards == 0) return NULL;
    LBAudioDetectiveDatabaseGroupRef g = calloc(1, sizeof *g);
    if (!g) return NULL;
    g->subfingerprintLength = inSubfingerprintLength; g->W = W;
    if (lbadcu_group_create(W, (inSubfingerprintLength + 1) / 2, inDevices, inNumberOfShards, &g->group) != LBAD_OK) { free(g); return NULL; }
    return g;
}

OSStatus LBAudioDetectiveDatabaseGroupDispose(LBAudioDetectiveDatabaseGroupRef g) {
    if (!g) return kLBAudioDetectiveArgumentInvalid;
    lbadcu_group_destroy(g->group); free(g);
    return noErr;
}

UInt32 LBAudioDetectiveDatabaseGroupGetNumberOfShards(LBAudioDetectiveDatabaseGroupRef g) { return g ? lbadcu_group_shards(g->group) : 0; }
UInt64 LBAudioDetectiveDatabaseGroupGetNumberOfClips(LBAudioDetectiveDatabaseGroupRef g) { return g ? lbadcu_group_clips(g->group) : 0; }
int LBAudioDetectiveDatabaseGroupGetShardDevice(LBAudioDetectiveDatabaseGroupRef g, UInt32 inShard) { return g ? lbadcu_group_shard_device(g->group, inShard) : -1; }
UInt32 LBAudioDetectiveDatabaseGroupGetShardNumberOfClips(LBAudioDetectiveDatabaseGroupRef g, UInt32 inShard) {
    lbadcu_db* db = g ? lbadcu_group_shard(g->group, inShard) : NULL;
    return db ? lbadcu_db_clips(db) : 0;
}

OSStatus LBAudioDetectiveDatabaseGroupAddPacked(LBAudioDetectiveDatabaseGroupRef g, const UInt32* inWords, UInt32 nClips, const UInt32* inCounts, UInt32 uniform, UInt64* outFirst) {
    if (!g || !inWords) return kLBAudioDetectiveArgumentInvalid;
    UInt64 first = lbadcu_group_next_id(g->group);
    OSStatus e = lbad_status(lbadcu_group_append(g->group, inWords, nClips, inCounts, uniform));
    if (e == noErr && outFirst) *outFirst = first;
    return e;
}

OSStatus LBAudioDetectiveDatabaseGroupAddFingerprint(LBAudioDetectiveDatabaseGroupRef g, LBAudioDetectiveFingerprintRef fp, UInt64* outClipIndex) {
    if (!g || !fp || fp->subfingerprintLength != g->subfingerprintLength) return kLBAudioDetectiveArgumentInvalid;
    UInt64 id = 0;
    OSStatus e = lbad_status(lbadcu_group_append_one(g->group, fp->words, fp->subfingerprintCount, &id));
    if (e == noErr && outClipIndex) *outClipIndex = id;
    return e;
}

OSStatus LBAudioDetectiveDatabaseGroupAddPackedDeviceToShard(LBAudioDetectiveDatabaseGroupRef g, UInt32 inShard, const UInt32* dWords, UInt32 nClips, UInt32 uniform, UInt64 inFirst, void* inProducerStream) {
    if (!g || !dWords) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_group_append_device(g->group, inShard, dWords, nClips, uniform, inFirst, inProducerStream));
}

static UInt32 group_pairs_for(LBAudioDetectiveDatabaseGroupRef g, UInt32 inRange) {
    if (inRange == 0) inRange = g->subfingerprintLength;                         /* LBAudioDetective.m:443-445 */
    return lbad_pairs_for_range(inRange, g->subfingerprintLength);
}

OSStatus LBAudioDetectiveDatabaseGroupSearchPacked(LBAudioDetectiveDatabaseGroupRef g, const UInt32* inQueryWords, UInt32 nQ, UInt32 qCount, UInt32 inRange, UInt32 inK,
                                                   Float32* outScores, UInt32* outClipIndices) {
    if (!g || (!inQueryWords && qCount) || !outScores || !outClipIndices) return kLBAudioDetectiveArgumentInvalid;
    UInt32 dummy[16] = {0};
    return lbad_status(lbadcu_group_search_host(g->group, qCount ? inQueryWords : dummy, nQ, qCount, group_pairs_for(g, inRange), inK, outScores, outClipIndices));
}

OSStatus LBAudioDetectiveDatabaseGroupSearch(LBAudioDetectiveDatabaseGroupRef g, const LBAudioDetectiveFingerprintRef* inQueries, UInt32 nQ, UInt32 inRange, UInt32 inK,
                                             Float32* outScores, UInt32* outClipIndices) {
    if (!g || !inQueries || nQ == 0) return kLBAudioDetectiveArgumentInvalid;
    UInt32 qCount = inQueries[0]->subfingerprintCount;
    for (UInt32 i = 0; i < nQ; i++)
        if (inQueries[i]->subfingerprintCount != qCount || inQueries[i]->subfingerprintLength != g->subfingerprintLength) return kLBAudioDetectiveArgumentInvalid;
    size_t per = (size_t)qCount * 2 * g->W;
    UInt32* words = malloc(((per * nQ) > 0 ? per * nQ : 1) * sizeof(UInt32));
    if (!words) return kLBAudioDetectiveArgumentInvalid;
    for (UInt32 i = 0; i < nQ; i++) memcpy(words + per * i, inQueries[i]->words, per * sizeof(UInt32));
    OSStatus e = LBAudioDetectiveDatabaseGroupSearchPacked(g, words, nQ, qCount, inRange, inK, outScores, outClipIndices);
    free(words);
    return e;
}

UInt64 LBAudioDetectiveDatabaseGroupGetKernelLaunchCount(LBAudioDetectiveDatabaseGroupRef g) { return g ? lbadcu_group_launches(g->group) : 0; }
Float64 LBAudioDetectiveDatabaseGroupGetLastSearchMilliseconds(LBAudioDetectiveDatabaseGroupRef g) { return g ? lbadcu_group_last_search_ms(g->group) : 0.0; }

/* ---- persistence: a packed binary file, so that a database can be reloaded without re-extracting ----
 * little-endian: char magic[8] = "LBADDB1\0"; u32 L; u32 W; u32 clips; u32 reserved; u64 subfingerprints;
 *                u32 counts[clips]; u32 words[subfingerprints][2*W]  (P plane then M plane per subfingerprint) */
static const char kMagic[8] = {'L', 'B', 'A', 'D', 'D', 'B', '1', '\0'};

OSStatus LBAudioDetectiveDatabaseSave(LBAudioDetectiveDatabaseRef d, const char* inPath) {
    if (!d || !inPath) return kLBAudioDetectiveArgumentInvalid;
    UInt32 clips = lbadcu_db_clips(d->db); UInt64 subfps = lbadcu_db_subfps(d->db);
    UInt32* counts = malloc(((size_t)clips ? clips : 1) * sizeof(UInt32));
    UInt32* words = malloc(((size_t)subfps ? subfps : 1) * 2 * d->W * sizeof(UInt32));
    if (!counts || !words) { free(counts); free(words); return kLBAudioDetectiveArgumentInvalid; }
    OSStatus e = lbad_status(lbadcu_db_download(d->db, words, counts));
    FILE* f = e == noErr ? fopen(inPath, "wb") : NULL;
    if (e == noErr && !f) e = kLBAudioDetectiveArgumentInvalid;
    if (f) {
        UInt32 hdr[4] = {d->subfingerprintLength, d->W, clips, 0};
        int ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(hdr, sizeof(UInt32), 4, f) == 4 && fwrite(&subfps, sizeof(UInt64), 1, f) == 1 &&
                 fwrite(counts, sizeof(UInt32), clips, f) == clips && fwrite(words, sizeof(UInt32), (size_t)subfps * 2 * d->W, f) == (size_t)subfps * 2 * d->W;
        if (fclose(f) != 0 || !ok) e = kLBAudioDetectiveArgumentInvalid;
    }
    free(counts); free(words);
    return e;
}

LBAudioDetectiveDatabaseRef LBAudioDetectiveDatabaseLoad(const char* inPath) {
    FILE* f = inPath ? fopen(inPath, "rb") : NULL;
    if (!f) return NULL;
    char magic[8]; UInt32 hdr[4]; UInt64 subfps = 0;
    LBAudioDetectiveDatabaseRef d = NULL; UInt32* counts = NULL; UInt32* words = NULL;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kMagic, 8) != 0 || fread(hdr, sizeof(UInt32), 4, f) != 4 || fread(&subfps, sizeof(UInt64), 1, f) != 1) goto done;
    if (lbad_words_per_plane(hdr[0]) != hdr[1] || hdr[1] == 0) goto done;
    counts = malloc(((size_t)hdr[2] ? hdr[2] : 1) * sizeof(UInt32));
    words = malloc(((size_t)subfps ? subfps : 1) * 2 * hdr[1] * sizeof(UInt32));
    if (!counts || !words) goto done;
    if (fread(counts, sizeof(UInt32), hdr[2], f) != hdr[2] || fread(words, sizeof(UInt32), (size_t)subfps * 2 * hdr[1], f) != (size_t)subfps * 2 * hdr[1]) goto done;
    UInt64 total = 0;
    for (UInt32 c = 0; c < hdr[2]; c++) total += counts[c];
    if (total != subfps) goto done;
    d = LBAudioDetectiveDatabaseNew(hdr[0]);
    if (d && hdr[2] && lbadcu_db_append(d->db, words, 0, hdr[2], counts, 0, NULL, -1) != LBAD_OK) { LBAudioDetectiveDatabaseDispose(d); d = NULL; }
done:
    fclose(f); free(counts); free(words);
    return d;
}
