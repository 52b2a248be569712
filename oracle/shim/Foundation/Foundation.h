/* TEST INFRASTRUCTURE — not part of the product.
 * Minimal Linux stand-in for <Foundation/Foundation.h>, just enough for the reference's
 * LBAudioDetective*.m files (read in place from /root/reference) to compile as C with
 * `gcc -x c -std=gnu99 -D__bridge=`.  Nothing here is shipped; see oracle/README.md. */
#ifndef LBAD_SHIM_FOUNDATION_H
#define LBAD_SHIM_FOUNDATION_H
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <ctype.h>

typedef uint8_t  UInt8;   typedef int8_t  SInt8;
typedef uint16_t UInt16;  typedef int16_t SInt16;
typedef uint32_t UInt32;  typedef int32_t SInt32;
typedef uint64_t UInt64;  typedef int64_t SInt64;
typedef float    Float32; typedef double  Float64;
typedef unsigned char Boolean;
typedef SInt32 OSStatus;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
enum { noErr = 0 };

/* Apple's NSObjCRuntime.h MIN/MAX: statement expressions, `a < b ? a : b` / `a < b ? b : a`.
 * The NaN behaviour of MAX (keeps the first operand when the second is NaN) is relied on by
 * LBAudioDetectiveFingerprint.m:144 when fp2 is empty (0/0). */
#define MIN(A,B) ({ __typeof__(A) __a = (A); __typeof__(B) __b = (B); __a < __b ? __a : __b; })
#define MAX(A,B) ({ __typeof__(A) __a = (A); __typeof__(B) __b = (B); __a < __b ? __b : __a; })

typedef struct LBADShimURL NSURL;          /* opaque; the memory-backed ExtAudioFile shim defines it */
typedef const struct LBADShimURL* CFURLRef;

static inline UInt32 CFSwapInt32HostToBig(UInt32 v) { return __builtin_bswap32(v); }
#endif
