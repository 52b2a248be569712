/* lbad_host.h — private declarations shared by the host-side C files of libLBAudioDetectiveCUDA. */
#ifndef LBAD_HOST_H
#define LBAD_HOST_H
#include "../../include/LBAudioDetective.h"
#include "../../include/LBAudioDetectiveDatabase.h"
#include "../../include/LBAudioDetectiveResample.h"
#include "lbad_cuda.h"

/* Fingerprint object (replaces struct LBAudioDetectiveFingerprint, LBAudioDetectiveFingerprint.m:10-14).
 * `booleans` is the reference's representation flattened: subfingerprint i at booleans + i*length.
 * `words` is the packed twin: subfingerprint i at words + i*2*W (0 words when the length is unsupported). */
struct LBAudioDetectiveFingerprint {
    UInt32 subfingerprintLength;
    UInt32 subfingerprintCount;
    UInt32 capacity;
    Boolean* booleans;
    UInt32* words;
};

UInt32 lbad_words_per_plane(UInt32 length);
void   lbad_pack_booleans(const Boolean* in, UInt32 length, UInt32 W, UInt32* out);
void   lbad_unpack_words(const UInt32* in, UInt32 length, UInt32 W, Boolean* out);
OSStatus lbad_status(int cuda_layer_code);
/* appends `count` packed subfingerprints, filling both representations */
OSStatus lbad_fingerprint_append_packed(LBAudioDetectiveFingerprintRef fp, const UInt32* words, UInt32 count);
UInt32 lbad_pairs_for_range(UInt32 range, UInt32 length);
#endif
