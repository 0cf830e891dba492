/* TEST INFRASTRUCTURE — FNV-1a-64 over a byte buffer (golden set hashes). */
#include <stddef.h>
#include <stdint.h>
uint64_t fnv1a64(const unsigned char *p, size_t n)
{
    uint64_t h = 0xCBF29CE484222325ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001B3ull; }
    return h;
}
