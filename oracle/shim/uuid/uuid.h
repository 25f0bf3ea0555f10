/* Test-infrastructure shim (NOT product code): the reference's metadata writer
 * includes <uuid/uuid.h> only to stamp a clip GUID into each sample
 * (EncoderSDK/SampleEncoder.cpp:26,763; MetadataWriter.cpp:27,333).  The dev
 * header is absent from this image, so oracle/_ref is built against this
 * header-only stand-in.  GUID bytes never touch the transform path. */
#pragma once
#include <stdlib.h>
typedef unsigned char uuid_t[16];
static inline void uuid_generate(uuid_t out) { for (int i = 0; i < 16; i++) out[i] = (unsigned char)rand(); }
