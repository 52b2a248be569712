/*
 * LBAudioDetective.h — the fingerprint extractor ("detective").
 *
 * Drop-in for the reference's LBAudioDetective/LBAudioDetective.h on the fingerprint path: same names,
 * argument order, OSStatus conventions and quirks.  h: = reference LBAudioDetective.h, m: = LBAudioDetective.m.
 * The two NSURL entry points (h:218, h:235) need AudioToolbox decode, which is out of scope; they are replaced
 * by PCM twins of otherwise identical shape (…ProcessPCM, …ComparePCM) taking float32 mono samples at the
 * processing sample rate.  All signal processing runs in hand-written CUDA kernels for sm_100a; there is no CPU
 * fallback (calls return kLBAudioDetectiveDeviceUnavailable when no usable device exists).
 *
 * Threading (as in the reference, which shares FFT scratch inside the struct, m:37-43): one detective per thread.
 * A detective is bound to the CUDA device that was current when it was created and owns one stream.
 */
#ifndef LBAUDIODETECTIVE_H
#define LBAUDIODETECTIVE_H
#include "LBAudioDetectiveTypes.h"
#include "LBAudioDetectiveFingerprint.h"
LBAD_EXTERN_C_BEGIN

/* h:14-20 / m:20-26.  kLBAudioDetectiveDefaultFingerprintComparisonRange is declared upstream (h:19) but never
 * defined; it is defined here as 0 ("use the detective's subfingerprint length", m:443-445). */
LBAD_API extern const OSStatus kLBAudioDetectiveArgumentInvalid;              /* 1 */
LBAD_API extern const UInt32 kLBAudioDetectiveDefaultWindowSize;              /* 2048 */
LBAD_API extern const UInt32 kLBAudioDetectiveDefaultAnalysisStride;          /* 64 */
LBAD_API extern const UInt32 kLBAudioDetectiveDefaultNumberOfPitchSteps;      /* 32 */
LBAD_API extern const UInt32 kLBAudioDetectiveDefaultNumberOfRowsPerFrame;    /* 128 (m:25; defined but undeclared upstream) */
LBAD_API extern const UInt32 kLBAudioDetectiveDefaultFingerprintComparisonRange; /* 0 */
LBAD_API extern const UInt32 kLBAudioDetectiveDefaultSubfingerprintLength;    /* 200 */
/* additions: CUDA-side failures, mapped into the OSStatus space */
LBAD_API extern const OSStatus kLBAudioDetectiveDeviceUnavailable;            /* -7001: no CUDA device / library built without kernels */
LBAD_API extern const OSStatus kLBAudioDetectiveDeviceError;                  /* -7002: a CUDA call failed (message on stderr) */

typedef struct LBAudioDetective *LBAudioDetectiveRef;

/* ---- reference surface ------------------------------------------------------------------------------ */

/* h:41, m:77-90. */
LBAD_API LBAudioDetectiveRef LBAudioDetectiveNew(void);
/* h:49, m:92-111.  Dispose(NULL) returns kLBAudioDetectiveArgumentInvalid, as upstream. */
LBAD_API OSStatus LBAudioDetectiveDispose(LBAudioDetectiveRef inDetective);
/* h:62, m:116-131: float32, packed, mono, 5512 Hz. */
LBAD_API AudioStreamBasicDescription LBAudioDetectiveDefaultProcessingFormat(void);
/* h:74, m:133-135 */
LBAD_API Float64 LBAudioDetectiveGetProcessingSampleRate(LBAudioDetectiveRef inDetective);
/* h:85, m:137-139 */
LBAD_API UInt32 LBAudioDetectiveGetNumberOfPitchSteps(LBAudioDetectiveRef inDetective);
/* h:96 and h:129 (declared twice upstream), m:141-143 */
LBAD_API UInt32 LBAudioDetectiveGetSubfingerprintLength(LBAudioDetectiveRef inDetective);
/* h:107, m:145-147 */
LBAD_API UInt32 LBAudioDetectiveGetWindowSize(LBAudioDetectiveRef inDetective);
/* h:118, m:149-151 */
LBAD_API UInt32 LBAudioDetectiveGetAnalysisStride(LBAudioDetectiveRef inDetective);
/* h:143: declared upstream but never defined (recording was removed).  Defined here: it sets the rate of the PCM handed to the
 * ...Recorded... entry points of LBAudioDetectiveResample.h (default 44100.0) and invalidates the cached resampler; rates that are
 * not positive return kLBAudioDetectiveArgumentInvalid. */
LBAD_API OSStatus LBAudioDetectiveSetRecordingSampleRate(LBAudioDetectiveRef inDetective, Float64 inSampleRate);
/* h:154, m:156-160 */
LBAD_API OSStatus LBAudioDetectiveSetProcessingSampleRate(LBAudioDetectiveRef inDetective, Float64 inSampleRate);
/* h:164, m:162-166 */
LBAD_API OSStatus LBAudioDetectiveSetNumberOfPitchSteps(LBAudioDetectiveRef inDetective, UInt32 inNumberOfPitchSteps);
/* h:174 and h:205, m:168-172 */
LBAD_API OSStatus LBAudioDetectiveSetSubfingerprintLength(LBAudioDetectiveRef inDetective, UInt32 inSubfingerprintLength);
/* h:184, m:174-195.  Keeps the upstream quirk: returns kLBAudioDetectiveArgumentInvalid when the
 * size IS a power of two, yet still applies it. */
LBAD_API OSStatus LBAudioDetectiveSetWindowSize(LBAudioDetectiveRef inDetective, UInt32 inWindowSize);
/* h:194, m:197-201 */
LBAD_API OSStatus LBAudioDetectiveSetAnalysisStride(LBAudioDetectiveRef inDetective, UInt32 inAnalysisStride);

/* Replaces LBAudioDetectiveProcessAudioURL (h:218, m:208-308).  inSamples: float32 mono PCM at the processing
 * sample rate (host memory).  *outFingerprint is a fresh fingerprint owned by the caller (m:297-300), with
 * ((n - window)/stride)/128 subfingerprints of GetSubfingerprintLength() Booleans.  Differences from upstream,
 * all where upstream is undefined: inNumberFrames < window returns kLBAudioDetectiveArgumentInvalid with an empty
 * fingerprint (upstream underflows, m:250); unsupported geometry (see LBAudioDetectiveCheckConfiguration)
 * returns kLBAudioDetectiveArgumentInvalid. */
LBAD_API OSStatus LBAudioDetectiveProcessPCM(LBAudioDetectiveRef inDetective, const Float32* inSamples, UInt64 inNumberFrames, LBAudioDetectiveFingerprintRef* outFingerprint);
/* Replaces LBAudioDetectiveCompareAudioURLs (h:235, m:442-464).  inComparisonRange 0 means the detective's
 * subfingerprint length (m:443-445); *outMatch is written only when no error occurred (m:456-458). */
LBAD_API OSStatus LBAudioDetectiveComparePCM(LBAudioDetectiveRef inDetective, const Float32* inSamples1, UInt64 inNumberFrames1, const Float32* inSamples2, UInt64 inNumberFrames2, UInt32 inComparisonRange, Float32* outMatch);

/* ---- additions: validation, batch extraction, stage dumps ------------------------------------------- */

/* noErr if the current (sample rate, window, stride, pitch steps, subfingerprint length) can be processed:
 * window a power of two in [256, 2048] whose band table stays inside the spectrum (4096 reads out
 * of bounds upstream), stride >= 1, pitch steps a power of two in [4, 64], even subfingerprint length in [2, 512]
 * and not more than 128*pitchSteps. */
LBAD_API OSStatus LBAudioDetectiveCheckConfiguration(LBAudioDetectiveRef inDetective);
/* Number of subfingerprints ProcessPCM yields for a clip of inNumberFrames samples (m:250-255); 0 if too short. */
LBAD_API UInt64 LBAudioDetectiveGetNumberOfSubfingerprintsForLength(LBAudioDetectiveRef inDetective, UInt64 inNumberFrames);
/* The band table the kernels use (m:361-383): outIndices[B+1], outLowBins[B], outHighBins[B]; any may be NULL. */
LBAD_API OSStatus LBAudioDetectiveGetBandTable(LBAudioDetectiveRef inDetective, UInt32* outIndices, UInt32* outLowBins, UInt32* outHighBins);

/* Batch extraction of inNumberOfClips equal-length clips laid out back to back with inClipStride samples between
 * clip starts (host memory; chunks are copied and processed on overlapping streams).  outWords receives
 * [clip][subfingerprint][2*W] packed words (see LBAudioDetectiveFingerprint.h), W = PackedWordsPerPlane(L). */
LBAD_API OSStatus LBAudioDetectiveProcessPCMBatch(LBAudioDetectiveRef inDetective, const Float32* inSamples, UInt32 inNumberOfClips, UInt64 inFramesPerClip, UInt64 inClipStride, UInt32* outWords);
/* Same for signed 16-bit PCM (the reference's recording format, essay p.VI ff.): half the bytes cross PCIe; samples are converted on
 * the device as x / 32768, exactly what a float32 client format would have delivered. */
LBAD_API OSStatus LBAudioDetectiveProcessPCMBatchInt16(LBAudioDetectiveRef inDetective, const SInt16* inSamples, UInt32 inNumberOfClips, UInt64 inFramesPerClip, UInt64 inClipStride, UInt32* outWords);
/* Which CUDA device a detective computes on.  By default (-1) it is the device that is current when the first computing call builds the
 * detective's device plan; a caller that does not use the CUDA runtime itself chooses one here (0 <= inDevice < number of devices;
 * kLBAudioDetectiveArgumentInvalid otherwise).  Changing the device drops the plan, like any other setter. */
LBAD_API OSStatus LBAudioDetectiveSetDevice(LBAudioDetectiveRef inDetective, int inDevice);
LBAD_API int LBAudioDetectiveGetDevice(LBAudioDetectiveRef inDetective);
/* LBAudioDetectiveProcessPCMBatch over SEVERAL detectives at once — normally one per GPU (LBAudioDetectiveSetDevice), all configured
 * alike: every detective's host thread runs the upload / kernel / download pipeline of LBAudioDetectiveProcessPCMBatch on the same
 * batch and takes its chunks (about 190 MB of PCM) from one shared cursor as its buffers drain — a GPU behind a slower host link ends
 * up with a smaller share —, nothing is exchanged between the GPUs (extraction shards by
 * clip).  outWords as above, in clip order; the result does not depend on the number of detectives or on who took what.  Returns the first error of any share; kLBAudioDetectiveArgumentInvalid if the configurations differ. */
LBAD_API OSStatus LBAudioDetectiveProcessPCMBatchSharded(const LBAudioDetectiveRef* inDetectives, UInt32 inNumberOfDetectives, const Float32* inSamples, UInt32 inNumberOfClips,
                                                         UInt64 inFramesPerClip, UInt64 inClipStride, UInt32* outWords);
/* Same, but inSamples and outWords are DEVICE pointers on the detective's device and the work is enqueued on
 * inStream (a cudaStream_t, NULL = the detective's own stream) without synchronising.  Calls on one detective share device scratch
 * (the spectral images between the two kernels): the library orders them on the device, so calls enqueued on different streams are
 * safe but run one after the other; use one detective per stream for concurrency. */
LBAD_API OSStatus LBAudioDetectiveProcessPCMBatchDevice(LBAudioDetectiveRef inDetective, const Float32* inDeviceSamples, UInt32 inNumberOfClips, UInt64 inFramesPerClip, UInt64 inClipStride, UInt32* outDeviceWords, void* inStream);
/* Stage dump for parity tests, one clip from host memory: outImages / outHaar are [subfp][128][B] floats (spectral
 * images before / after the Haar transform), outBooleans is [subfp][L]; any may be NULL.  inUseFusedKernel selects
 * the register-FFT fast path (only valid for window 2048 / 32 pitch steps) or the generic shared-memory-FFT path. */
LBAD_API OSStatus LBAudioDetectiveProcessPCMStages(LBAudioDetectiveRef inDetective, const Float32* inSamples, UInt64 inNumberFrames, Float32* outImages, Float32* outHaar, Boolean* outBooleans, Boolean inUseFusedKernel);
/* Stage dump of a batch (layout of LBAudioDetectiveProcessPCMBatch): outWords [clip][subfp][2*W] (required), outImages / outHaar
 * [clip][subfp][128][B] (either may be NULL).  The same kernels as the batch entry point, chunk by chunk. */
LBAD_API OSStatus LBAudioDetectiveProcessPCMBatchStages(LBAudioDetectiveRef inDetective, const Float32* inSamples, UInt32 inNumberOfClips, UInt64 inFramesPerClip, UInt64 inClipStride,
                                                       UInt32* outWords, Float32* outImages, Float32* outHaar, Boolean inUseFusedKernel);
/* Haar transform (Frame.m:113-153) + ordered top-t sign extraction (Frame.m:165-191) of inCount host images
 * [128][B]; outHaar [inCount][128][B] and outBooleans [inCount][L] may be NULL. */
LBAD_API OSStatus LBAudioDetectiveTransformImages(LBAudioDetectiveRef inDetective, const Float32* inImages, UInt32 inCount, Float32* outHaar, Boolean* outBooleans);
/* ---- streaming extraction (addition): the recording use case of the essay (p.23-24) without the whole-clip requirement ---- */
typedef struct LBAudioDetectiveStream *LBAudioDetectiveStreamRef;
/* Borrows inDetective, which must outlive the stream and keep its configuration (window, stride, pitch steps, subfingerprint length,
 * processing sample rate): an append after any of them changed returns kLBAudioDetectiveArgumentInvalid.  NULL if the configuration
 * is unsupported. */
LBAD_API LBAudioDetectiveStreamRef LBAudioDetectiveStreamNew(LBAudioDetectiveRef inDetective);
LBAD_API OSStatus LBAudioDetectiveStreamDispose(LBAudioDetectiveStreamRef inStream);
/* Appends PCM; every frame that the one-shot LBAudioDetectiveProcessPCM would produce for the samples appended so far is
 * extracted (on the GPU) and added to the stream's fingerprint — after any sequence of appends the fingerprint equals the one-shot
 * result on the concatenation. */
LBAD_API OSStatus LBAudioDetectiveStreamAppend(LBAudioDetectiveStreamRef inStream, const Float32* inSamples, UInt64 inNumberFrames);
/* The growing fingerprint, owned by the stream (copy it to keep it past Dispose). */
LBAD_API LBAudioDetectiveFingerprintRef LBAudioDetectiveStreamGetFingerprint(LBAudioDetectiveStreamRef inStream);
/* Samples buffered and not yet covered by an emitted subfingerprint. */
LBAD_API UInt64 LBAudioDetectiveStreamGetNumberOfPendingFrames(LBAudioDetectiveStreamRef inStream);

/* Kernels launched by this detective since creation (for bench.py's gpu_launches). */
LBAD_API UInt64 LBAudioDetectiveGetKernelLaunchCount(LBAudioDetectiveRef inDetective);
/* Total device time in ms of the dominant extraction kernel (FFT + band energies) over the launches since the last
 * call with inReset != 0 (CUDA events on the launching stream); returns the number of launches measured. */
LBAD_API UInt32 LBAudioDetectiveGetKernelTiming(LBAudioDetectiveRef inDetective, Boolean inEnable, Boolean inReset, Float64* outTotalMilliseconds);
/* Same for the second kernel of the path (Haar transform + ordered top-t + packing). */
LBAD_API UInt32 LBAudioDetectiveGetTransformKernelTiming(LBAudioDetectiveRef inDetective, Boolean inEnable, Boolean inReset, Float64* outTotalMilliseconds);

LBAD_EXTERN_C_END
#endif
