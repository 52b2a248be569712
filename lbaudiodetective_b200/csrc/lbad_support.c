/*
 * lbad_support.c — exported bench/test support entry points that are NOT part of the reference surface
 * (device-side synthetic PCM, random codes, microbenchmarks).  Declared in include/LBAudioDetectiveSupport.h.
 */
#include "lbad_host.h"
#include "../../include/LBAudioDetectiveSupport.h"

OSStatus LBAudioDetectiveSupportSynthesizeDevice(Float32* d_out, UInt32 nClips, UInt64 clipLen, UInt64 clipStride, UInt64 firstClipId, UInt64 baseSeed, Float64 sampleRate, void* stream) {
    return lbad_status(lbadcu_synth_device(d_out, nClips, clipLen, clipStride, firstClipId, baseSeed, sampleRate, stream));
}
OSStatus LBAudioDetectiveSupportRandomCodesDevice(UInt32* d_words, UInt64 nSubfps, UInt32 subfingerprintLength, UInt64 seed, void* stream) {
    UInt32 W = lbad_words_per_plane(subfingerprintLength);
    if (!W) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_random_codes_device(d_words, nSubfps, W, (subfingerprintLength + 1) / 2, seed, 0, stream));
}
OSStatus LBAudioDetectiveSupportRandomCodesDeviceAt(UInt32* d_words, UInt64 nSubfps, UInt32 subfingerprintLength, UInt64 seed, UInt64 firstSubfp, void* stream) {
    UInt32 W = lbad_words_per_plane(subfingerprintLength);
    if (!W) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_random_codes_device(d_words, nSubfps, W, (subfingerprintLength + 1) / 2, seed, firstSubfp, stream));
}
int LBAudioDetectiveSupportDeviceCount(void) { return lbadcu_device_count(); }
OSStatus LBAudioDetectiveSupportMicrobench(Float64* outFp32Tflops, Float64* outPopcGops, Float64* outLop3Gops) {
    return lbad_status(lbadcu_microbench(outFp32Tflops, outPopcGops, outLop3Gops));
}
const char* LBAudioDetectiveSupportLastError(void) { return lbadcu_last_error(); }
Boolean LBAudioDetectiveSupportDeviceAvailable(void) { return lbadcu_device_available() == LBAD_OK; }
