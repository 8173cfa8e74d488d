"""tinyfusers_b200/synthetic.py (what example/sd1.py and bench.py's GPU arm build the model from) against the oracle's own copy of
the generators: the product side and the checker must start from bit-identical weights, inputs, schedule and FLOP counts."""
import torch


def test_generators_match_the_oracle_bit_for_bit(oracle):
    from tinyfusers_b200 import synthetic as S
    a, b = {}, {}
    oracle.add_res_block(a, "rb", 320, 640, seed=1234)
    S.add_res_block(b, "rb", 320, 640, seed=1234)
    oracle.add_spatial_transformer(a, "st", 320, 768, seed=99)
    S.add_spatial_transformer(b, "st", 320, 768, seed=99)
    oracle.add_attn_block(a, "ab", 64, seed=7)
    S.add_attn_block(b, "ab", 64, seed=7)
    a.update(oracle.make_clip_state_dict(layers=1))
    b.update(S.make_clip_state_dict(layers=1))
    assert sorted(a) == sorted(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    for x, y in zip(oracle.make_inputs(2, 16, seed=5, ctx_seed=6), S.make_inputs(2, 16, seed=5, ctx_seed=6)):
        assert torch.equal(x, y)


def test_structure_tables_schedule_and_flops_match(oracle):
    from tinyfusers_b200 import synthetic as S
    assert S.UNET_INPUT_BLOCKS == oracle.UNET_INPUT_BLOCKS and S.UNET_MIDDLE_BLOCK == oracle.UNET_MIDDLE_BLOCK
    assert S.UNET_OUTPUT_BLOCKS == oracle.UNET_OUTPUT_BLOCKS and S.VAE_DECODER_SZ == oracle.VAE_DECODER_SZ
    for steps in (3, 50):
        ta, aa, pa = oracle.sampler_schedule(steps)
        tb, ab, pb = S.sampler_schedule(steps)
        assert ta == tb and torch.equal(aa, ab) and torch.equal(pa, pb)
    for n, hw in ((2, 64), (16, 64), (8, 96)):
        assert S.unet_step_flops(n, hw, hw) == oracle.unet_step_flops(n, hw, hw)
    assert abs(S.unet_step_flops(2, 64, 64) / 1e9 - 1606.5) < 0.1


def test_product_entry_points_do_not_import_the_oracle():
    """example/sd1.py and bench.py's GPU arm build everything from tinyfusers_b200.synthetic; `oracle` may only appear in bench.py's
    cpu_baseline / --impl reference legs."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert "oracle" not in open(os.path.join(root, "example", "sd1.py")).read()
    for dirpath, _, files in os.walk(os.path.join(root, "tinyfusers_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), os.path.join(dirpath, f)
    bench = open(os.path.join(root, "bench.py")).read()
    ours = bench[bench.index("def run_ours("):bench.index("def main(")]
    assert ours.count("from oracle import") == 1 and ours.index("from oracle import") > ours.index("CPU baseline")
