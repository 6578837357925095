"""Generates tests/golden/*.json from sources independent of this repository's C++/CUDA code:
python hashlib (BLAKE2b), the EIP-152 text (vectors 4-7, transcribed), the reference's own
vector 5 literal (/root/reference/blake2f-circuit/src/blake2f.rs:193-247), big-int field
arithmetic, and a pure-Python transcription of rand_xorshift 0.3.0.

Run from the repo root: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))

H = "48c9bdf267e6096a3ba7ca8485ae67bb2bf894fe72f36e3cf1361d5f3af54fa5d182e6ad7f520e511f6c3e2b8c68059b6bbd41fbabd9831f79217e1319cde05b"
M = "616263" + "00" * 125
T = "0300000000000000" + "0000000000000000"


def rec(rounds, f):
    return "%08x" % rounds + H + M + T + "%02x" % f


EIP152 = [
    # EIP-152 test vectors 4..7 (vector 5 is also the reference's INPUTS_OUTPUTS literal)
    {"name": "eip152-4", "input": rec(0, 1),
     "output": "08c9bcf367e6096a3ba7ca8485ae67bb2bf894fe72f36e3cf1361d5f3af54fa5d282e6ad7f520e511f6c3e2b8c68059b9442be0454267ce079217e1319cde05b"},
    {"name": "eip152-5", "input": rec(12, 1),
     "output": "ba80a53f981c4d0d6a2797b69f12f6e94c212f14685ac4b74b12bb6fdbffa2d17d87c5392aab792dc252d5de4533cc9518d38aa8dbf1925ab92386edd4009923"},
    {"name": "eip152-6", "input": rec(12, 0),
     "output": "75ab69d3190a562c51aef8d88f1c2775876944407270c42c9844252c26d2875298743e7f6d5ea2f2d3e8d226039cd31b4e426ac4f2d3d666a610c2116fde4735"},
    {"name": "eip152-7", "input": rec(1, 1),
     "output": "b63a380cb2897d521994a85234ee2c181b5f844d2c624c002677e9703449d2fba551b3a8333bcdf5f2f7e08993d53923de3d64fcc68c034e717b9293fed7a421"},
]

IV = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1,
      0x510e527fade682d1, 0x9b05688c2b3e6c1f, 0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]


def hashlib_single_block(msg):
    """F input/outputs for a single-block unkeyed BLAKE2b-512 hash, output from hashlib."""
    assert len(msg) <= 128
    h = list(IV)
    h[0] ^= 0x01010040
    hb = struct.pack("<8Q", *h).hex()
    m = (msg + bytes(128 - len(msg))).hex()
    t = struct.pack("<2Q", len(msg), 0).hex()
    return {"name": "hashlib-%d" % len(msg), "input": "%08x" % 12 + hb + m + t + "01",
            "output": hashlib.blake2b(msg).hexdigest()}


def xorshift(seed, n):
    x, y, z, w = struct.unpack("<4I", seed)
    out = []
    for _ in range(n):
        vals = []
        for _ in range(2):
            t = (x ^ (x << 11)) & 0xFFFFFFFF
            x, y, z = y, z, w
            w = (w ^ (w >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF
            vals.append(w)
        out.append(vals[0] | (vals[1] << 32))
    return out


def main():
    vecs = list(EIP152)
    for msg in [b"", b"abc", bytes(range(128)), b"zk-odst" * 9]:
        vecs.append(hashlib_single_block(msg))
    json.dump(vecs, open(os.path.join(HERE, "eip152.json"), "w"), indent=1)

    seed = bytes([0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54,
                  0x06, 0xbc, 0xe5])
    json.dump({"seed": seed.hex(), "next_u64": [hex(v) for v in xorshift(seed, 32)]},
              open(os.path.join(HERE, "xorshift.json"), "w"), indent=1)

    # spread-table spot rows from the reference's own test (spread_table.rs:684-723)
    spots = [[0, 0b000, 0b000000], [0, 0b001, 0b000001], [0, 0b010, 0b000100], [0, 0b011, 0b000101],
             [0, 0b100, 0b010000], [0, 0b101, 0b010001], [0, 0xFF, 0x5555], [1, 0x100, 0x10000],
             [1, 0x7FFF, 0x15555555], [2, 0x8000, 0x40000000], [2, 0xFFFF, 0x55555555]]
    json.dump(spots, open(os.path.join(HERE, "spread_spots.json"), "w"))

    # field constants (the pasta_curves literals) and big-int KATs
    p = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
    q = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001
    fields = {}
    for name, mod in (("fp", p), ("fq", q)):
        g = 5
        fields[name] = {
            "MOD": hex(mod), "R": hex(pow(2, 256, mod)), "R2": hex(pow(2, 512, mod)),
            "R3": hex(pow(2, 768, mod)), "INV": hex((-pow(mod, -1, 1 << 64)) % (1 << 64)),
            "ROOT_OF_UNITY": hex(pow(g, (mod - 1) >> 32, mod)), "DELTA": hex(pow(g, 1 << 32, mod)),
            "ZETA": hex(pow(pow(g, (mod - 1) // 3, mod), 2, mod)),
        }
    fields["literals"] = {  # as recalled from pasta_curves 0.5.1 in SURVEY.md Appendix A.1
        "fp_ROOT_OF_UNITY": "0x2bce74deac30ebda362120830561f81aea322bf2b7bb7584bdad6fabd87ea32f",
        "fp_DELTA": "0x0a757d0f0006ab6cbd455b7112a5049df5e4f3f13eee56366a6ccd20dd7b9ba2",
        "fp_ZETA": "0x12ccca834acdba712caad5dc57aab1b01d1f8bd237ad31491dad5ebdfdfe4ab9",
        "fq_ROOT_OF_UNITY": "0x2de6a9b8746d3f589e5c4dfd492ae26e9bb97ea3c106f049a70e2c1102b6d05f",
        "fp_INV": "0x992d30ecffffffff", "fq_INV": "0x8c46eb20ffffffff",
    }
    json.dump(fields, open(os.path.join(HERE, "fields.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
