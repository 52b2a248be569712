/*
 * LBAudioDetectiveResample.h — front end (addition): recording-rate PCM -> processing-rate PCM on the GPU.
 *
 * In the reference this step is not code of its own: LBAudioDetectiveProcessAudioURL sets the processing format
 * (5512 Hz mono float32, LBAudioDetective.m:116-131) as the ExtAudioFile client format (m:229) and Apple's converter
 * resamples the 44.1 kHz file while it is read (m:275).  AudioToolbox is closed and out of scope, so the conversion is
 * DEFINED here (it cannot be pinned to Apple's converter; parity for this step is against oracle/lbad_oracle.c only):
 *
 *   rho = inRate / outRate >= 1,  D = max(1, floor(rho / 2)),  rho2 = rho / D  (in [1, 4)).
 *   stage 1 (skipped when D = 1): y1[n] = sum_{t < 2 H1 + 1} g[t] x[D n + t - H1],  H1 = 6 D,
 *             g[t] = sinc((t - H1) / D) / D * kaiser_8((t - H1) / (H1 + 1)), normalised to sum 1   (x is 0 outside the clip)
 *   stage 2:  out[m] = S0 + a (S1 - S0),  S_r = sum_{i < 2 H2} Hc[p + r][i] y1[i0 + i - H2 + 1],
 *             m rho2 = i0 + phi (i0 integer, 0 <= phi < 1),  p = floor(64 phi),  a = 64 phi - p,
 *             Hc[p][i] = gamma sinc(gamma u) kaiser_9(u / H2) for |u| < H2 (else 0),  u = i - H2 + 1 - p / 64,
 *             gamma = 0.97 / rho2,  H2 = ceil(10 / gamma);  every row normalised to sum 1
 *   tables in double rounded once to float32; sums are float32 FMA chains — stage 2 in increasing tap order, stage 1 in polyphase order
 *   (t = p, p + D, p + 2 D, ... for p = 0 .. D - 1, the order in which a de-interleaved input is consumed); m rho2 in double;
 *   outFrames = floor(inFrames * outRate / inRate + 1e-9).
 * For 44.1 kHz -> 5512 Hz: D = 4 (49 taps), rho2 = 2.00018, 42 taps.  The cutoff sits at 0.97 of the output Nyquist because the band
 * table only reads 231-2043 Hz: the response is flat there (error <= 1.2e-3 at 2043 Hz) and everything that could alias
 * below 2043 Hz (input above 5512 - 2043 = 3469 Hz) is down by 88 dB or more (tests/test_resample.py).
 */
#ifndef LBAUDIODETECTIVE_RESAMPLE_H
#define LBAUDIODETECTIVE_RESAMPLE_H
#include "LBAudioDetective.h"
LBAD_EXTERN_C_BEGIN

/* h:143 declares LBAudioDetectiveSetRecordingSampleRate and the reference never defines it; here it sets the rate of the PCM
 * handed to the ...Recorded... entry points below (default 44100.0).  Rates below the processing rate are rejected. */
LBAD_API Float64 LBAudioDetectiveGetRecordingSampleRate(LBAudioDetectiveRef inDetective);
/* Frames LBAudioDetectiveResamplePCM yields for inNumberFrames recorded frames. */
LBAD_API UInt64 LBAudioDetectiveGetResampledLength(LBAudioDetectiveRef inDetective, UInt64 inNumberFrames);
/* Host PCM at the recording rate -> host PCM at the processing rate (outSamples holds GetResampledLength(inNumberFrames) frames). */
LBAD_API OSStatus LBAudioDetectiveResamplePCM(LBAudioDetectiveRef inDetective, const Float32* inSamples, UInt64 inNumberFrames, Float32* outSamples);
/* LBAudioDetectiveProcessPCM on recorded-rate PCM: resampled and fingerprinted on the device without a host round trip. */
LBAD_API OSStatus LBAudioDetectiveProcessRecordedPCM(LBAudioDetectiveRef inDetective, const Float32* inSamples, UInt64 inNumberFrames, LBAudioDetectiveFingerprintRef* outFingerprint);
/* Batch form over DEVICE buffers: inNumberOfClips clips of inFramesPerClip recorded frames, inClipStride apart; outDeviceWords as in
 * LBAudioDetectiveProcessPCMBatchDevice for clips of GetResampledLength(inFramesPerClip) frames.  Enqueued on inStream. */
LBAD_API OSStatus LBAudioDetectiveProcessRecordedPCMBatchDevice(LBAudioDetectiveRef inDetective, const Float32* inDeviceSamples, UInt32 inNumberOfClips, UInt64 inFramesPerClip, UInt64 inClipStride, UInt32* outDeviceWords, void* inStream);

LBAD_EXTERN_C_END
#endif
