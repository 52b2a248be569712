/*
 * LBAudioDetectiveSupport.h — bench/test support exported by libLBAudioDetectiveCUDA.so.  None of this exists in
 * the reference; it is here so that bench.py can build full-size inputs on the device and measure the pipe rates
 * the kernels are bound by.  Pointers named d_* are device pointers; `stream` is a cudaStream_t (NULL = default).
 */
#ifndef LBAUDIODETECTIVE_SUPPORT_H
#define LBAUDIODETECTIVE_SUPPORT_H
#include "LBAudioDetectiveTypes.h"
LBAD_EXTERN_C_BEGIN
/* chirp + tone + uniform noise per SURVEY.md §8(d); clip c (id firstClipId + c) is written at d_out + c*clipStride */
LBAD_API OSStatus LBAudioDetectiveSupportSynthesizeDevice(Float32* d_out, UInt32 nClips, UInt64 clipLen, UInt64 clipStride, UInt64 firstClipId, UInt64 baseSeed, Float64 sampleRate, void* stream);
/* random rank-sign codes (one sign bit per rank) as packed subfingerprints, for search timing */
LBAD_API OSStatus LBAudioDetectiveSupportRandomCodesDevice(UInt32* d_words, UInt64 nSubfps, UInt32 subfingerprintLength, UInt64 seed, void* stream);
/* measured FP32 FMA rate (TFLOP/s) and POPC / LOP3 lane-operation rates (Gop/s) of the current device */
LBAD_API OSStatus LBAudioDetectiveSupportMicrobench(Float64* outFp32Tflops, Float64* outPopcGops, Float64* outLop3Gops);
LBAD_API const char* LBAudioDetectiveSupportLastError(void);
LBAD_API Boolean LBAudioDetectiveSupportDeviceAvailable(void);
/* number of CUDA devices visible to this process (0 if none) */
LBAD_API int LBAudioDetectiveSupportDeviceCount(void);
/* LBAudioDetectiveSupportRandomCodesDevice for a slice of a larger set: subfingerprint i of the call is subfingerprint firstSubfp + i of
 * the whole (the code depends on seed and the GLOBAL index only, so a sharded database holds the same clips however it is cut) */
LBAD_API OSStatus LBAudioDetectiveSupportRandomCodesDeviceAt(UInt32* d_words, UInt64 nSubfps, UInt32 subfingerprintLength, UInt64 seed, UInt64 firstSubfp, void* stream);
LBAD_EXTERN_C_END
#endif
