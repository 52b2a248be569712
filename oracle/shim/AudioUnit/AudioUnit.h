/* TEST INFRASTRUCTURE — empty stand-in; the reference imports but never uses AudioUnit (LBAudioDetective.m:10). */
