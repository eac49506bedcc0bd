"""Host-side logic of the drop-in that does not need a GPU."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from vq_gan_b200 import VectorQuantizer
from vq_gan_b200 import distributed as vdist


def test_constructor_matches_reference_contract():
    g = load_golden("init_seed42_k128_d256")
    torch.manual_seed(42)
    vq = VectorQuantizer(128, 256, 0.25)
    # same codebook AND same RNG position afterwards as the reference constructor
    np.testing.assert_array_equal(vq.embedding.weight.detach().numpy(), g["weight"])
    np.testing.assert_array_equal(torch.rand(4).numpy(), g["next_rand"])
    assert sorted(vq.state_dict().keys()) == list(g["state_keys"]) == ["embedding.weight"]
    assert (vq.num_embeddings, vq.embedding_dim, vq.commitment_cost) == (128, 256, 0.25)
    assert isinstance(vq.embedding, torch.nn.Embedding)


def test_reference_checkpoint_loads_strictly():
    vq = VectorQuantizer(16, 8)
    sd = {"embedding.weight": torch.randn(16, 8)}
    vq.load_state_dict(sd, strict=True)
    assert torch.equal(vq.embedding.weight.detach(), sd["embedding.weight"])


def test_cpu_inputs_are_rejected_not_emulated():
    vq = VectorQuantizer(16, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vq(torch.randn(1, 4, 2, 2))
    with pytest.raises(RuntimeError):
        vq.get_codebook_entry(torch.zeros(1, 2, 2, dtype=torch.int64))
    with pytest.raises(RuntimeError):
        vq.get_codebook_usage(torch.zeros(1, 2, 2, dtype=torch.int64))


def test_shape_errors_like_reference():
    vq = VectorQuantizer(16, 4)
    with pytest.raises(RuntimeError):
        vq(torch.randn(1, 5, 2, 2))  # reference: view(-1, 4) then matmul shape error
    with pytest.raises(RuntimeError):
        vq(torch.randn(4, 2, 2))
    with pytest.raises(ValueError):
        VectorQuantizer(16, 4, return_format="nope")


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 64, 65536, 50000):
        for world in (1, 2, 3, 8):
            spans = [vdist.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_stats_pack_roundtrip_is_exact_for_large_counts():
    dE = torch.randn(5, 3)
    hist = torch.tensor([0, 1, 4095, 4096, 1 << 20, (1 << 31) + 12345], dtype=torch.int64)
    scal = torch.tensor([3.5, -1.0])
    flat = vdist.pack_stats(dE, hist, scal)
    a, s, h = vdist.unpack_stats(flat * 1.0, dE.shape, 2, hist.numel())
    assert torch.equal(a, dE) and torch.equal(s, scal) and torch.equal(h, hist)
    # summing R copies stays exact (what the all-reduce does)
    a, s, h = vdist.unpack_stats(flat * 8.0, dE.shape, 2, hist.numel())
    assert torch.equal(h, hist * 8)


def test_quant_conv_is_checkpoint_compatible_with_nn_conv2d():
    """`pre_quant_conv` / `post_quant_conv` (vq_vae.py:74-79) swap: same parameters and state_dict keys as
    nn.Conv2d(cin, cout, 1); no CPU path."""
    from vq_gan_b200 import QuantConv1x1
    ref = torch.nn.Conv2d(8, 16, kernel_size=1)
    mine = QuantConv1x1(8, 16)
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    assert all(mine.state_dict()[k].shape == v.shape for k, v in ref.state_dict().items())
    mine.load_state_dict(ref.state_dict(), strict=True)
    # the reference's own call forms (vq_vae.py:75-76: nn.Conv2d(z_channels, embedding_dim, kernel_size=1))
    for m in (QuantConv1x1(8, 16, 1), QuantConv1x1(8, 16, kernel_size=1), QuantConv1x1(8, 16, (1, 1), bias=True)):
        assert m.bias is not None and m.weight.shape == (16, 8, 1, 1)
    assert QuantConv1x1(8, 16, 1, bias=False).bias is None
    for bad in (dict(kernel_size=3), dict(stride=2), dict(padding=1)):
        with pytest.raises(ValueError):
            QuantConv1x1(8, 16, **bad)
    with pytest.raises(RuntimeError):
        mine(torch.randn(1, 8, 4, 4))
    with pytest.raises(RuntimeError):
        mine(torch.randn(8, 4, 4))


def test_extension_modules_host_contract():
    from vq_gan_b200 import EMAVectorQuantizer
    from vq_gan_b200.graphs import GraphedVectorQuantizer
    torch.manual_seed(42)
    ema = EMAVectorQuantizer(32, 8, 0.25, decay=0.9)
    assert sorted(ema.state_dict().keys()) == ["cluster_size", "embed_sum", "embedding.weight"]
    assert not ema.embedding.weight.requires_grad
    # same RNG contract as the reference constructor: the codebook equals the plain module's under one seed
    torch.manual_seed(42)
    from vq_gan_b200 import VectorQuantizer
    plain = VectorQuantizer(32, 8, 0.25)
    assert torch.equal(plain.embedding.weight.detach(), ema.embedding.weight.detach())
    # a reference checkpoint loads with strict=False (only the two EMA buffers are missing)
    missing = ema.load_state_dict(plain.state_dict(), strict=False)
    assert sorted(missing.missing_keys) == ["cluster_size", "embed_sum"] and not missing.unexpected_keys
    with pytest.raises(RuntimeError):
        GraphedVectorQuantizer(plain, torch.randn(1, 8, 4, 4))


def test_sharded_argmin_key_orders_like_aten_argmin():
    """(score, index) keys: signed MIN picks the smallest score, lowest index on ties; NaN is minimal;
    -0.0 ties with +0.0 (distributed.py:90)."""
    from oracle import vq_oracle as orc
    d = torch.tensor([1.5, -2.0, float("nan"), 0.0, -0.0, float("inf"), -float("inf")])
    idx = torch.arange(7)
    keys = orc.argmin_key(d, idx)
    order = torch.argsort(keys).tolist()
    assert order[0] == 2                      # NaN first
    assert order[1] == 6                      # then -inf
    assert order[2] == 1
    assert order[3:5] == [3, 4]               # +0.0 / -0.0 tie: lower index first
    assert order[5:] == [0, 5]
    # two NaN scores on different shards: the lower global index wins
    k = orc.argmin_key(torch.tensor([float("nan"), float("nan")]), torch.tensor([700, 12]))
    assert int(k.min() & 0xFFFFFFFF) == 12


def test_loss_dict_is_a_dict_of_floats_on_every_read_path():
    """`LossDict`: the logged floats of quantizer.py:106-107 are fetched at first use; whatever way the dict is read,
    the caller sees what the reference returns (a dict with one tensor and two Python floats)."""
    import copy
    import pickle
    from vq_gan_b200 import LossDict

    def fresh():
        return LossDict(torch.tensor(1.25), torch.tensor(0.5))   # CPU tensors stand in for the device scalars

    want = {"vq_loss": torch.tensor(1.25), "codebook_loss": 0.5, "commitment_loss": 0.5}
    d = fresh()
    assert isinstance(d, dict) and list(d) == ["vq_loss", "codebook_loss", "commitment_loss"] and len(d) == 3
    assert d._pending is not None                      # nothing read yet: no sync happened in "forward"
    assert d["vq_loss"].item() == 1.25 and d._pending is not None
    assert d["codebook_loss"] == 0.5 and isinstance(d["codebook_loss"], float) and d._pending is None
    for read in (lambda x: dict(x), lambda x: {**x}, lambda x: x.copy(), lambda x: dict(x.items()),
                 lambda x: copy.copy(x), lambda x: pickle.loads(pickle.dumps(x)), lambda x: {k: x.get(k) for k in x}):
        got = read(fresh())
        assert type(got["commitment_loss"]) is float and got["commitment_loss"] == 0.5, read
        assert set(got) == set(want)
    assert list(fresh().values())[1:] == [0.5, 0.5]
    plain = {}
    plain.update(fresh())
    assert plain["codebook_loss"] == 0.5
    assert "0.5" in repr(fresh())
    # VQVAE.forward adds a key (vq_vae.py:158); user code may overwrite the logged ones
    d = fresh()
    d["codebook_usage_ratio"] = 0.75
    assert d["codebook_usage_ratio"] == 0.75 and d["commitment_loss"] == 0.5
    d = fresh()
    d["codebook_loss"] = 9.0
    assert d["codebook_loss"] == 9.0 and d["commitment_loss"] == 0.5


def test_round2_ops_reject_cpu_tensors():
    """No CPU path exists for the layers next to the quantizer either: host tensors raise instead of being emulated."""
    from vq_gan_b200 import GroupNormSiLU, QuantConv1x1, ops
    x = torch.randn(2, 32, 8, 8)
    gy = torch.randn(2, 16, 8, 8)
    with pytest.raises(RuntimeError):
        ops.conv1x1_param_grads(gy, x)
    with pytest.raises(RuntimeError):
        ops.conv1x1(x, torch.randn(16, 32), None)
    with pytest.raises(RuntimeError):
        ops.groupnorm_silu(x, torch.ones(32), torch.zeros(32), 4, 1e-6)
    with pytest.raises(RuntimeError):
        QuantConv1x1(32, 16)(x)
    with pytest.raises(RuntimeError):
        GroupNormSiLU(4, 32)(x)
