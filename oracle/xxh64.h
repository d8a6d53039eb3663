/*
 * ORACLE (test infrastructure, not product code).
 *
 * XXH64, written from the published xxHash specification (xxHash r39 / XXH64).
 *
 * Why it is here: Rapid orders every ring and computes configuration ids with
 * net.openhft:zero-allocation-hashing:0.8 `LongHashFunction.xx(seed)`
 * (rapid/pom.xml:79-83; call sites rapid/src/main/java/com/vrg/rapid/MembershipView.java:47,
 * :548-553, :568, :580-581).  That jar is NOT vendored under /root/reference and there is no JVM
 * in this image, so the algorithm is restated from the public XXH64 spec:
 *   hashBytes(ByteBuffer) == XXH64(bytes[position..limit), seed)
 *   hashInt(int v)        == XXH64(4 little-endian bytes of v, seed)
 *   hashLong(long v)      == XXH64(8 little-endian bytes of v, seed)
 * Pinned against the canonical XXH64 vectors, python `xxhash` 3.7.0 and libxxhash 0.8.2 in
 * tests/test_oracle_xxh64.py.  The hashInt/hashLong == "XXH64 of the LE bytes" reading cannot be
 * confirmed without the jar: RING-HASH PARITY IS UNPINNED; java/PinRingHash.java
 * prints the values to diff wherever a JVM + the 0.8 jar exist.
 */
#ifndef RAPID_ORACLE_XXH64_H
#define RAPID_ORACLE_XXH64_H

#include <stdint.h>
#include <stddef.h>
#include <string.h>

static const uint64_t ORC_P1 = 0x9E3779B185EBCA87ULL;
static const uint64_t ORC_P2 = 0xC2B2AE3D27D4EB4FULL;
static const uint64_t ORC_P3 = 0x165667B19E3779F9ULL;
static const uint64_t ORC_P4 = 0x85EBCA77C2B2AE63ULL;
static const uint64_t ORC_P5 = 0x27D4EB2F165667C5ULL;

static inline uint64_t orc_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t orc_read64le(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}

static inline uint32_t orc_read32le(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

static inline uint64_t orc_xxh64_round(uint64_t acc, uint64_t input) {
    acc += input * ORC_P2;
    acc = orc_rotl64(acc, 31);
    return acc * ORC_P1;
}

static inline uint64_t orc_xxh64_merge(uint64_t acc, uint64_t val) {
    val = orc_xxh64_round(0, val);
    acc ^= val;
    return acc * ORC_P1 + ORC_P4;
}

static inline uint64_t orc_xxh64(const void* data, size_t len, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* const end = p + len;
    uint64_t h;
    if (len >= 32) {
        const uint8_t* const limit = end - 32;
        uint64_t v1 = seed + ORC_P1 + ORC_P2;
        uint64_t v2 = seed + ORC_P2;
        uint64_t v3 = seed;
        uint64_t v4 = seed - ORC_P1;
        do {
            v1 = orc_xxh64_round(v1, orc_read64le(p)); p += 8;
            v2 = orc_xxh64_round(v2, orc_read64le(p)); p += 8;
            v3 = orc_xxh64_round(v3, orc_read64le(p)); p += 8;
            v4 = orc_xxh64_round(v4, orc_read64le(p)); p += 8;
        } while (p <= limit);
        h = orc_rotl64(v1, 1) + orc_rotl64(v2, 7) + orc_rotl64(v3, 12) + orc_rotl64(v4, 18);
        h = orc_xxh64_merge(h, v1);
        h = orc_xxh64_merge(h, v2);
        h = orc_xxh64_merge(h, v3);
        h = orc_xxh64_merge(h, v4);
    } else {
        h = seed + ORC_P5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) {
        const uint64_t k1 = orc_xxh64_round(0, orc_read64le(p));
        h ^= k1;
        h = orc_rotl64(h, 27) * ORC_P1 + ORC_P4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)orc_read32le(p) * ORC_P1;
        h = orc_rotl64(h, 23) * ORC_P2 + ORC_P3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * ORC_P5;
        h = orc_rotl64(h, 11) * ORC_P1;
        ++p;
    }
    h ^= h >> 33;
    h *= ORC_P2;
    h ^= h >> 29;
    h *= ORC_P3;
    h ^= h >> 32;
    return h;
}

/* LongHashFunction.xx(seed).hashInt(v): XXH64 of the 4 little-endian bytes. */
static inline uint64_t orc_xx_hash_int(int32_t v, uint64_t seed) {
    uint8_t b[4];
    const uint32_t u = (uint32_t)v;
    b[0] = (uint8_t)u; b[1] = (uint8_t)(u >> 8); b[2] = (uint8_t)(u >> 16); b[3] = (uint8_t)(u >> 24);
    return orc_xxh64(b, 4, seed);
}

/* LongHashFunction.xx(seed).hashLong(v): XXH64 of the 8 little-endian bytes. */
static inline uint64_t orc_xx_hash_long(int64_t v, uint64_t seed) {
    uint8_t b[8];
    uint64_t u = (uint64_t)v;
    for (int i = 0; i < 8; ++i) { b[i] = (uint8_t)u; u >>= 8; }
    return orc_xxh64(b, 8, seed);
}

#endif
