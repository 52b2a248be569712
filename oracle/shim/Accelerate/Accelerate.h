/* TEST INFRASTRUCTURE — not part of the product.  Stand-in for <Accelerate/Accelerate.h>:
 * only the five vDSP entry points the reference calls (LBAudioDetective.m:106,179,192,353-355). */
#ifndef LBAD_SHIM_ACCELERATE_H
#define LBAD_SHIM_ACCELERATE_H
#include <Foundation/Foundation.h>
typedef struct DSPComplex { float real; float imag; } DSPComplex, COMPLEX;
typedef struct DSPSplitComplex { float* realp; float* imagp; } DSPSplitComplex, COMPLEX_SPLIT;
typedef struct LBADShimFFTSetup* FFTSetup;
typedef unsigned long vDSP_Length; typedef long vDSP_Stride;
typedef int FFTDirection; typedef int FFTRadix;
enum { FFT_RADIX2 = 0 };
enum { FFT_FORWARD = +1, FFT_INVERSE = -1 };
FFTSetup vDSP_create_fftsetup(vDSP_Length log2n, FFTRadix radix);
void vDSP_destroy_fftsetup(FFTSetup setup);
void vDSP_ctoz(const DSPComplex* C, vDSP_Stride IC, const DSPSplitComplex* Z, vDSP_Stride IZ, vDSP_Length N);
void vDSP_ztoc(const DSPSplitComplex* Z, vDSP_Stride IZ, DSPComplex* C, vDSP_Stride IC, vDSP_Length N);
void vDSP_fft_zrip(FFTSetup setup, const DSPSplitComplex* C, vDSP_Stride IC, vDSP_Length log2n, FFTDirection dir);
#endif
