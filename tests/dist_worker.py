"""One rank of the multi-GPU MSM-split proof (launched by tests/test_dist_gpu.py and bench.py)."""
import sys


def main():
    rank, world, uid_hex, k, ncomp, out_path = (int(sys.argv[1]), int(sys.argv[2]), sys.argv[3],
                                                 int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import zk_odst_b200 as zk
    ctx = zk.Context(rank)
    ctx.dist_init(bytes.fromhex(uid_hex), rank, world)
    assert ctx.dist_info() == (rank, world)
    ctx.params_generate_substitute(k, zk.REFERENCE_SEED)
    ctx.keygen(12, ncomp)
    inputs = zk.synthetic_inputs(ncomp)
    proof = ctx.create_proof(inputs, ncomp, zk.REFERENCE_SEED)
    with open(out_path + ".%d" % rank, "wb") as f:
        f.write(ctx.vk_bytes() + proof)
    ctx.close()


if __name__ == "__main__":
    main()
