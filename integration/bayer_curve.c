/* bayer_curve.c -- the encoder's default Bayer encode curve, exactly as Codec/frame.c:5208-5222 builds it inside
 * ConvertBYR4ToFrame16s (log base 90 over 1 << 14 input levels, 12-bit output).  A C translation unit on purpose: the
 * reference's own macro / inline function (Common/AVIExtendedHeader.h:153 lin2log) calls log10() on float arguments, which
 * is the double function in C (as in frame.c) but the float overload in C++ -- a handful of table entries differ by one.
 * Part of the reference-side binding (the host builds the curve table today and hands it to cfb_codec_set_bayer_curve). */
#include <math.h>
#include <stdint.h>
#include "AVIExtendedHeader.h"

void cfhd_shim_default_bayer_curve(uint16_t *table /* 1 << 14 entries */)
{
    const int max_value = 1 << 14, precision = 12;
    int i;
    table[0] = 0;
    for (i = 1; i < max_value; i++)
        table[i] = (uint16_t)(int)(CURVE_LIN2LOG((float)i / (float)max_value, 90) * (float)((1 << precision) - 1));
}
