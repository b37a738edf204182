/* test_swar.c -- exhaustive check (all 2^32 words) of the two SWAR identities the scan kernels rely on
 * (sliceslice_rs_b200/csrc/ss_device.cuh):
 *   swar_zero_term(x)  = (x - 0x01010101) & ~x          & 0x80808080 != 0   <=>  x has a zero byte
 *   swar_zero_exact(x) = ~(((x & 0x7F7F7F7F) + 0x7F7F7F7F) | x | 0x7F7F7F7F)  ==  0x80 in exactly the zero bytes
 * The first may set extra bits above a true zero byte but never fires without one, so no candidate is
 * missed and none is invented; the second is the exact per-position mask used by the hit path. */
#include <stdint.h>
#include <stdio.h>

static inline uint32_t swar_zero_term(uint32_t x) { return (x - 0x01010101u) & ~x; }
static inline uint32_t swar_zero_exact(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu); }

int main(void)
{
    uint64_t bad = 0;
    for (uint64_t v = 0; v <= 0xFFFFFFFFull; v++) {
        const uint32_t x = (uint32_t)v;
        uint32_t ref = 0;
        for (int b = 0; b < 4; b++)
            if (((x >> (8 * b)) & 0xFF) == 0)
                ref |= 0x80u << (8 * b);
        const int any = (swar_zero_term(x) & 0x80808080u) != 0;
        if (any != (ref != 0) || swar_zero_exact(x) != ref) {
            if (bad++ < 5)
                printf("mismatch at %08x\n", x);
        }
    }
    printf("%s: 4294967296 words, %llu mismatches\n", bad ? "FAILED" : "ok", (unsigned long long)bad);
    return bad ? 1 : 0;
}
