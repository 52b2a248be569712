/*
 * LBAudioDetectiveDatabase.h — batched form of the reference's matcher (an addition; the reference has no such type).
 *
 * The reference compares one pair at a time with LBAudioDetectiveFingerprintCompareToFingerprint (FP.h:134,
 * FP.m:119-149); its only "database" is the linear scan + argmax of the essay's server.  This type
 * keeps many fingerprints packed on one GPU and evaluates
 *        score[q][c] = LBAudioDetectiveFingerprintCompareToFingerprint(clip_c, query_q, range)
 * (archive first, query second — the argument order of LBAudioDetectiveTests.m:68) for whole query batches in one
 * __popc kernel, bit-exact with the pairwise function, and returns per-query top-k ordered (score desc, clip asc).
 * Multi-GPU: LBAudioDetectiveDatabaseGroup (below) shards a database over the GPUs of one process; with one process per GPU, each holds a
 * shard (LBAudioDetectiveDatabaseSetClipIndexBase) and LBAudioDetectiveDatabaseMergeTopK[Device] merges the gathered shard results.
 * Searches on one database are serialised on the device whatever streams they are enqueued on (they share the partial-list buffers).
 */
#ifndef LBAUDIODETECTIVE_DATABASE_H
#define LBAUDIODETECTIVE_DATABASE_H
#include "LBAudioDetectiveTypes.h"
#include "LBAudioDetectiveFingerprint.h"
LBAD_EXTERN_C_BEGIN

typedef struct LBAudioDetectiveDatabase *LBAudioDetectiveDatabaseRef;

/* All fingerprints in a database share one subfingerprint length L.  Bound to the current CUDA device. */
LBAD_API LBAudioDetectiveDatabaseRef LBAudioDetectiveDatabaseNew(UInt32 inSubfingerprintLength);
LBAD_API OSStatus LBAudioDetectiveDatabaseDispose(LBAudioDetectiveDatabaseRef inDatabase);
LBAD_API UInt32 LBAudioDetectiveDatabaseGetNumberOfClips(LBAudioDetectiveDatabaseRef inDatabase);
LBAD_API UInt64 LBAudioDetectiveDatabaseGetNumberOfSubfingerprints(LBAudioDetectiveDatabaseRef inDatabase);
/* Global clip id of local clip 0 (for sharded databases; default 0). */
LBAD_API OSStatus LBAudioDetectiveDatabaseSetClipIndexBase(LBAudioDetectiveDatabaseRef inDatabase, UInt32 inBase);
/* Appends one fingerprint; *outClipIndex (optional) receives its local clip index. */
LBAD_API OSStatus LBAudioDetectiveDatabaseAddFingerprint(LBAudioDetectiveDatabaseRef inDatabase, LBAudioDetectiveFingerprintRef inFingerprint, UInt32* outClipIndex);
/* Appends inNumberOfClips clips given packed (host memory): inWords is the concatenation of all their
 * subfingerprints (2*W words each); inCounts[c] subfingerprints per clip, or NULL with inUniformCount each. */
LBAD_API OSStatus LBAudioDetectiveDatabaseAddPacked(LBAudioDetectiveDatabaseRef inDatabase, const UInt32* inWords, UInt32 inNumberOfClips, const UInt32* inCounts, UInt32 inUniformCount);
/* Same with inDeviceWords already on the device (uniform counts only): device-to-device append.  inProducerStream is the cudaStream_t
 * the words are being written on (e.g. the stream given to LBAudioDetectiveProcessPCMBatchDevice): the append is ordered after the work
 * enqueued there so far.  NULL: the words are complete already (the caller has synchronised).  Returns when the clips are searchable. */
LBAD_API OSStatus LBAudioDetectiveDatabaseAddPackedDevice(LBAudioDetectiveDatabaseRef inDatabase, const UInt32* inDeviceWords, UInt32 inNumberOfClips, UInt32 inUniformCount, void* inProducerStream);

/* Top-k search.  Queries: inNumberOfQueries packed fingerprints with inQueryCount subfingerprints each (host
 * memory, [query][subfp][2*W]).  inRange 0 means L.  outScores/outClipIndices: [query][k], ordered by score
 * descending then clip index ascending; unused slots (fewer than k clips) hold score -1 and index 0xFFFFFFFF.
 * outAllScores (optional, host) receives the full [query][clip] score matrix — for parity tests on small inputs. */
LBAD_API OSStatus LBAudioDetectiveDatabaseSearchPacked(LBAudioDetectiveDatabaseRef inDatabase, const UInt32* inQueryWords, UInt32 inNumberOfQueries, UInt32 inQueryCount,
                                                      UInt32 inRange, UInt32 inK, Float32* outScores, UInt32* outClipIndices, Float32* outAllScores);
/* Convenience over fingerprint objects (all with the same number of subfingerprints). */
LBAD_API OSStatus LBAudioDetectiveDatabaseSearch(LBAudioDetectiveDatabaseRef inDatabase, const LBAudioDetectiveFingerprintRef* inQueries, UInt32 inNumberOfQueries,
                                                UInt32 inRange, UInt32 inK, Float32* outScores, UInt32* outClipIndices);
/* Device-resident form used for timing: query words and outputs are device pointers, enqueued on inStream
 * (cudaStream_t; NULL = the database's stream), no synchronisation. */
LBAD_API OSStatus LBAudioDetectiveDatabaseSearchDevice(LBAudioDetectiveDatabaseRef inDatabase, const UInt32* inDeviceQueryWords, UInt32 inNumberOfQueries, UInt32 inQueryCount,
                                                      UInt32 inRange, UInt32 inK, Float32* outDeviceScores, UInt32* outDeviceClipIndices, void* inStream);
/* Merges inNumberOfLists top-k lists per query ([list][query][k], e.g. one per GPU after a gather) into one,
 * ordered (score desc, clip index asc) — the result equals a single-GPU search over the union.  Host memory. */
LBAD_API OSStatus LBAudioDetectiveDatabaseMergeTopK(const Float32* inScores, const UInt32* inClipIndices, UInt32 inNumberOfLists, UInt32 inNumberOfQueries, UInt32 inK,
                                                   Float32* outScores, UInt32* outClipIndices);
/* Same with every buffer on the device (e.g. straight out of an NCCL all-gather), enqueued on inStream without synchronising. */
LBAD_API OSStatus LBAudioDetectiveDatabaseMergeTopKDevice(const Float32* inDeviceScores, const UInt32* inDeviceClipIndices, UInt32 inNumberOfLists, UInt32 inNumberOfQueries, UInt32 inK,
                                                         Float32* outDeviceScores, UInt32* outDeviceClipIndices, void* inStream);
/* Same for lists that are not back to back: list l starts inListStride ELEMENTS after list l - 1 in both arrays — e.g. the receive
 * buffer of ONE all-gather whose per-rank payload is [scores | indices] (inDeviceClipIndices = inDeviceScores + nQ*k reinterpreted,
 * inListStride = 2*nQ*k): the merge reads the gather buffer in place, no repacking copy. */
LBAD_API OSStatus LBAudioDetectiveDatabaseMergeTopKDeviceStrided(const Float32* inDeviceScores, const UInt32* inDeviceClipIndices, UInt32 inNumberOfLists, UInt64 inListStride,
                                                                UInt32 inNumberOfQueries, UInt32 inK, Float32* outDeviceScores, UInt32* outDeviceClipIndices, void* inStream);

/* ---- a database sharded over several GPUs, driven by one process ------------------------------------------------
 * The reference's matcher scans one archive at a time on one core (FP.m:119-149, called in a loop by LBAudioDetectiveTests.m:64-70);
 * its callers are C programs in one process.  A group gives such a caller the multi-GPU form of the same scan without bringing a
 * transport of its own: shard i lives on CUDA device inDevices[i] (a device may hold several shards), clips are numbered globally in
 * the order they are added, a search runs the per-shard top-k kernels concurrently, moves the [query][k] lists to the first shard's
 * device by peer copies (NVLink where the devices are peers) and merges them there.  The result equals what ONE database holding all
 * the clips returns, bit for bit: scores as LBAudioDetectiveFingerprintCompareToFingerprint(clip, query, range), order (score
 * descending, global clip index ascending). */
typedef struct LBAudioDetectiveDatabaseGroup *LBAudioDetectiveDatabaseGroupRef;
LBAD_API LBAudioDetectiveDatabaseGroupRef LBAudioDetectiveDatabaseGroupNew(UInt32 inSubfingerprintLength, const int* inDevices, UInt32 inNumberOfShards);
LBAD_API OSStatus LBAudioDetectiveDatabaseGroupDispose(LBAudioDetectiveDatabaseGroupRef inGroup);
LBAD_API UInt32 LBAudioDetectiveDatabaseGroupGetNumberOfShards(LBAudioDetectiveDatabaseGroupRef inGroup);
LBAD_API UInt64 LBAudioDetectiveDatabaseGroupGetNumberOfClips(LBAudioDetectiveDatabaseGroupRef inGroup);
/* CUDA device of shard inShard, clips it holds. */
LBAD_API int    LBAudioDetectiveDatabaseGroupGetShardDevice(LBAudioDetectiveDatabaseGroupRef inGroup, UInt32 inShard);
LBAD_API UInt32 LBAudioDetectiveDatabaseGroupGetShardNumberOfClips(LBAudioDetectiveDatabaseGroupRef inGroup, UInt32 inShard);
/* Appends clips given packed in host memory (layout of LBAudioDetectiveDatabaseAddPacked); they get the next global clip indices
 * (*outFirstClipIndex, optional, receives the first) and are spread over the shards in contiguous blocks. */
LBAD_API OSStatus LBAudioDetectiveDatabaseGroupAddPacked(LBAudioDetectiveDatabaseGroupRef inGroup, const UInt32* inWords, UInt32 inNumberOfClips, const UInt32* inCounts, UInt32 inUniformCount, UInt64* outFirstClipIndex);
LBAD_API OSStatus LBAudioDetectiveDatabaseGroupAddFingerprint(LBAudioDetectiveDatabaseGroupRef inGroup, LBAudioDetectiveFingerprintRef inFingerprint, UInt64* outClipIndex);
/* Appends clips whose words already sit on the device of shard inShard (uniform counts), with the global clip indices
 * [inFirstClipIndex, inFirstClipIndex + inNumberOfClips), which must lie above every index the shard holds already.
 * inProducerStream as in LBAudioDetectiveDatabaseAddPackedDevice. */
LBAD_API OSStatus LBAudioDetectiveDatabaseGroupAddPackedDeviceToShard(LBAudioDetectiveDatabaseGroupRef inGroup, UInt32 inShard, const UInt32* inDeviceWords, UInt32 inNumberOfClips, UInt32 inUniformCount,
                                                                     UInt64 inFirstClipIndex, void* inProducerStream);
/* Top-k over all shards; arguments and results as LBAudioDetectiveDatabaseSearchPacked (host memory), clip indices global. */
LBAD_API OSStatus LBAudioDetectiveDatabaseGroupSearchPacked(LBAudioDetectiveDatabaseGroupRef inGroup, const UInt32* inQueryWords, UInt32 inNumberOfQueries, UInt32 inQueryCount,
                                                           UInt32 inRange, UInt32 inK, Float32* outScores, UInt32* outClipIndices);
LBAD_API OSStatus LBAudioDetectiveDatabaseGroupSearch(LBAudioDetectiveDatabaseGroupRef inGroup, const LBAudioDetectiveFingerprintRef* inQueries, UInt32 inNumberOfQueries,
                                                     UInt32 inRange, UInt32 inK, Float32* outScores, UInt32* outClipIndices);
LBAD_API UInt64 LBAudioDetectiveDatabaseGroupGetKernelLaunchCount(LBAudioDetectiveDatabaseGroupRef inGroup);
/* Device time in ms of the last search as the first shard's stream saw it (query upload to merged result on the host). */
LBAD_API Float64 LBAudioDetectiveDatabaseGroupGetLastSearchMilliseconds(LBAudioDetectiveDatabaseGroupRef inGroup);

/* Persistence: packed binary file (header: L, W, clip count, per-clip subfingerprint counts; body: the bit planes), so that a
 * database is reloaded without re-extracting.  Load returns NULL on a missing / malformed file or without a CUDA device. */
LBAD_API OSStatus LBAudioDetectiveDatabaseSave(LBAudioDetectiveDatabaseRef inDatabase, const char* inPath);
LBAD_API LBAudioDetectiveDatabaseRef LBAudioDetectiveDatabaseLoad(const char* inPath);
/* Number of CompareSubfingerprints evaluations one search performs (for compares/s). */
LBAD_API UInt64 LBAudioDetectiveDatabaseComparesPerQuery(LBAudioDetectiveDatabaseRef inDatabase, UInt32 inQueryCount);
LBAD_API UInt64 LBAudioDetectiveDatabaseGetKernelLaunchCount(LBAudioDetectiveDatabaseRef inDatabase);
/* Device time of the search kernel alone (see LBAudioDetectiveGetKernelTiming). */
LBAD_API UInt32 LBAudioDetectiveDatabaseGetKernelTiming(LBAudioDetectiveDatabaseRef inDatabase, Boolean inEnable, Boolean inReset, Float64* outTotalMilliseconds);

LBAD_EXTERN_C_END
#endif
