/*
 * LBAudioDetectiveTypes.h — Linux stand-ins for the CoreFoundation/CoreAudio scalar types that appear in the
 * reference's public headers (LBAudioDetective.h:9-10 imports Foundation + AudioToolbox for them).
 * Same widths and meanings as on iOS, so the signatures below are source-compatible with callers of the reference.
 */
#ifndef LBAUDIODETECTIVE_TYPES_H
#define LBAUDIODETECTIVE_TYPES_H
#include <stdint.h>
#include <stddef.h>

#if !defined(__MACTYPES__) && !defined(LBAD_HAVE_APPLE_TYPES)
typedef uint8_t  UInt8;
typedef int16_t  SInt16;
typedef uint32_t UInt32;
typedef int32_t  SInt32;
typedef uint64_t UInt64;
typedef int64_t  SInt64;
typedef float    Float32;
typedef double   Float64;
typedef unsigned char Boolean;
typedef SInt32   OSStatus;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#ifndef noErr
#define noErr 0
#endif

/* Layout-compatible with CoreAudio's AudioStreamBasicDescription (40 bytes), returned by value from
 * LBAudioDetectiveDefaultProcessingFormat() exactly as the reference does (LBAudioDetective.h:62). */
typedef struct AudioStreamBasicDescription {
    Float64 mSampleRate;
    UInt32  mFormatID;
    UInt32  mFormatFlags;
    UInt32  mBytesPerPacket;
    UInt32  mFramesPerPacket;
    UInt32  mBytesPerFrame;
    UInt32  mChannelsPerFrame;
    UInt32  mBitsPerChannel;
    UInt32  mReserved;
} AudioStreamBasicDescription;
#define kAudioFormatLinearPCM     0x6c70636du   /* 'lpcm' */
#define kAudioFormatFlagIsFloat   (1u << 0)
#define kAudioFormatFlagIsPacked  (1u << 3)
#endif

#ifdef __cplusplus
#define LBAD_EXTERN_C_BEGIN extern "C" {
#define LBAD_EXTERN_C_END }
#else
#define LBAD_EXTERN_C_BEGIN
#define LBAD_EXTERN_C_END
#endif
#define LBAD_API __attribute__((visibility("default")))

#endif
