"""EIP-152-shaped inputs (host-side data preparation; not on the proving path).

Synthetic workload per BASELINE.md §3: rounds = 12, `h, m, t` = consecutive little-endian u64
outputs of an XorShiftRng seeded with the reference harness's seed
(/root/reference benchmarking/src/blake2f_circuit_bench.rs:41-44), `f = i & 1`; the stream index
is XORed into the last seed byte.
"""
import struct

REFERENCE_SEED = bytes(
    [0x59, 0x62, 0xBE, 0x5D, 0x76, 0x3D, 0x31, 0x8D, 0x17, 0xDB, 0x37, 0x32, 0x54, 0x06, 0xBC, 0xE5]
)


class XorShiftRng:
    """rand_xorshift 0.3.0 XorShiftRng (xorshift128, 32-bit words)."""

    def __init__(self, seed=REFERENCE_SEED):
        s = list(struct.unpack("<4I", seed))
        if not any(s):
            s = [0x0BAD5EED] * 4
        self.x, self.y, self.z, self.w = s

    def next_u32(self):
        t = (self.x ^ (self.x << 11)) & 0xFFFFFFFF
        self.x, self.y, self.z = self.y, self.z, self.w
        self.w = (self.w ^ (self.w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
        return self.w

    def next_u64(self):
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)


def eip152_record(rounds, h, m, t, f):
    """213-byte precompile input: rounds u32 BE | h 8xu64 LE | m 16xu64 LE | t 2xu64 LE | f."""
    assert len(h) == 8 and len(m) == 16 and len(t) == 2
    return struct.pack(">I", rounds) + struct.pack("<8Q", *h) + struct.pack("<16Q", *m) + \
        struct.pack("<2Q", *t) + bytes([f])


def synthetic_inputs(n, rounds=12, stream=0, seed=REFERENCE_SEED):
    seed = seed[:15] + bytes([seed[15] ^ (stream & 0xFF)])
    rng = XorShiftRng(seed)
    out = bytearray()
    for i in range(n):
        h = [rng.next_u64() for _ in range(8)]
        m = [rng.next_u64() for _ in range(16)]
        t = [rng.next_u64() for _ in range(2)]
        out += eip152_record(rounds, h, m, t, i & 1)
    return bytes(out)
