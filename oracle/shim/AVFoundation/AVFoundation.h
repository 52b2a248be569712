/* TEST INFRASTRUCTURE — empty stand-in; the reference imports but never uses AVFoundation (LBAudioDetective.m:9). */
