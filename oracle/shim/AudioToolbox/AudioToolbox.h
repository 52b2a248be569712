/* TEST INFRASTRUCTURE — not part of the product.  Stand-in for <AudioToolbox/AudioToolbox.h>. */
#ifndef LBAD_SHIM_AUDIOTOOLBOX_H
#define LBAD_SHIM_AUDIOTOOLBOX_H
#include <Foundation/Foundation.h>

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

typedef struct AudioBuffer { UInt32 mNumberChannels; UInt32 mDataByteSize; void* mData; } AudioBuffer;
typedef struct AudioBufferList { UInt32 mNumberBuffers; AudioBuffer mBuffers[1]; } AudioBufferList;

enum { kAudioFormatLinearPCM = 0x6c70636d /* 'lpcm' */ };
enum { kAudioFormatFlagIsFloat = 1u << 0, kAudioFormatFlagIsPacked = 1u << 3 };
enum { kExtAudioFileProperty_ClientDataFormat = 0x63666d74 /* 'cfmt' */,
       kExtAudioFileProperty_FileLengthFrames = 0x2366726d /* '#frm' */ };

typedef struct LBADShimExtAudioFile* ExtAudioFileRef;
typedef struct LBADShimAudioConverter* AudioConverterRef;

OSStatus ExtAudioFileOpenURL(CFURLRef inURL, ExtAudioFileRef* outFile);
OSStatus ExtAudioFileDispose(ExtAudioFileRef inFile);
OSStatus ExtAudioFileSetProperty(ExtAudioFileRef inFile, UInt32 inID, UInt32 inSize, const void* inData);
OSStatus ExtAudioFileGetProperty(ExtAudioFileRef inFile, UInt32 inID, UInt32* ioSize, void* outData);
OSStatus ExtAudioFileRead(ExtAudioFileRef inFile, UInt32* ioNumberFrames, AudioBufferList* ioData);
OSStatus ExtAudioFileSeek(ExtAudioFileRef inFile, SInt64 inFrameOffset);

OSStatus AudioConverterNew(const AudioStreamBasicDescription* inFrom, const AudioStreamBasicDescription* inTo, AudioConverterRef* out);
OSStatus AudioConverterConvertComplexBuffer(AudioConverterRef c, UInt32 inNumberFrames, const AudioBufferList* in, AudioBufferList* out);
OSStatus AudioConverterDispose(AudioConverterRef c);
#endif
