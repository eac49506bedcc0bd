"""Pruned exact tier of the fp16 tensor search (csrc/vqb_search_pruned.cu): COLLAPSED codebooks -- many codes a rounding
error apart, so that no low-precision pass can certify a winner -- must give exactly what the plain fp32 search gives
(same FMA chain: indices AND minimum scores bit-identical, lowest index on exact ties), whether the tier takes the
list of uncertified tokens (`stats[3]` = its length) or declines (nothing to prune)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _collapsed(K, D, n_centres, spread, tokens, tok_sigma, seed):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(n_centres, D, generator=g)
    E = centres[torch.randint(0, n_centres, (K,), generator=g)] + spread * torch.randn(K, D, generator=g)
    z = centres[torch.randint(0, n_centres, (tokens,), generator=g)] + tok_sigma * torch.randn(tokens, D, generator=g)
    return z, E


def _as_images(rows, HW):
    N, D = rows.shape
    B = N // HW
    return rows[:B * HW].view(B, HW, D).permute(0, 2, 1).contiguous().view(B, D, HW, 1)


@pytest.mark.parametrize("K,D,centres,tokens,HW,expect_tier", [
    (4096, 64, 8, 16384, 1024, True),     # the tier takes every token
    (2500, 100, 5, 16384, 256, True),     # K not a multiple of the 128-code tiles, D not a multiple of 32
    (16384, 256, 16, 16384, 1024, True),  # C3's codebook size
    (4096, 64, 1, 16384, 1024, False),    # ONE centre: every tile survives for every token, the tier declines
    (4096, 64, 8, 8192, 1024, False),     # a small batch (< 16384 tokens) stays with the plain list search
])
def test_collapsed_codebook_equals_the_fp32_search(K, D, centres, tokens, HW, expect_tier):
    from vq_gan_b200 import ops
    z, E = _collapsed(K, D, centres, 1e-4, tokens, 0.05, seed=K + D)
    zc, Ec = _as_images(z, HW).cuda(), E.cuda()
    idx, dmin, st = ops.search(zc, Ec, 4)
    ref_idx, ref_dmin, _ = ops.search(zc, Ec, 2)
    st = st.tolist()
    assert st[1] == 4
    assert st[0] >= 0.9 * tokens  # nothing certifies on such a codebook
    assert (st[3] == st[0]) if expect_tier else (st[3] == 0), st
    assert torch.equal(idx, ref_idx)
    assert torch.equal(dmin, ref_dmin)


def test_exact_duplicates_keep_the_lowest_index():
    """Whole clusters of IDENTICAL codes: every score ties exactly, the original (unsorted) index decides.  The prepare
    pass hides the later copies from the tensor pass (codebook_shadow_kernel), so the tokens certify again."""
    from vq_gan_b200 import ops
    g = torch.Generator().manual_seed(2)
    centres = torch.randn(6, 64, generator=g)
    E = centres[torch.randint(0, 6, (4096,), generator=g)].contiguous()
    z = centres[torch.randint(0, 6, (16384,), generator=g)] + 0.05 * torch.randn(16384, 64, generator=g)
    zc, Ec = _as_images(z, 1024).cuda(), E.cuda()
    idx, dmin, st = ops.search(zc, Ec, 4)
    ref_idx, ref_dmin, _ = ops.search(zc, Ec, 2)
    assert st.tolist()[0] < 0.01 * 16384, st.tolist()   # (before the shadowing: every token uncertified)
    # certified tokens get their score from the candidate re-score (another, equally exact fp32 summation order)
    assert torch.equal(idx, ref_idx) and torch.allclose(dmin, ref_dmin, rtol=1e-5, atol=1e-4)
    # the winner is the FIRST code of the token's cluster
    first = torch.stack([(E == E[i]).all(dim=1).nonzero()[0, 0] for i in ref_idx.reshape(-1)[:64].cpu()])
    assert torch.equal(first, ref_idx.reshape(-1)[:64].cpu())


@pytest.mark.parametrize("copies,D", [(4, 64), (8, 256), (64, 128)])
def test_codebook_of_exact_copies_certifies(copies, D):
    """A healthy codebook restarted by COPYING codes (every distinct code `copies` times, shuffled): four or more equal
    scores used to fill every candidate slot of the tensor pass and send each token to the exact full search."""
    from vq_gan_b200 import ops
    g = torch.Generator().manual_seed(copies + D)
    K = 8192
    base = torch.randn(K // copies, D, generator=g)
    E = base.repeat_interleave(copies, dim=0)[torch.randperm(K, generator=g)].contiguous()
    z = torch.randn(16384, D, generator=g)
    zc, Ec = _as_images(z, 1024).cuda(), E.cuda()
    idx, dmin, st = ops.search(zc, Ec, 4)
    ref_idx, ref_dmin, _ = ops.search(zc, Ec, 2)
    assert st.tolist()[0] < 0.01 * 16384, st.tolist()
    assert torch.equal(idx, ref_idx) and torch.allclose(dmin, ref_dmin, rtol=1e-5, atol=1e-4)
    # every winner is the first of its copies
    rows = E[ref_idx.reshape(-1)[:32].cpu()]
    first = torch.stack([(E == r).all(dim=1).nonzero()[0, 0] for r in rows])
    assert torch.equal(first, ref_idx.reshape(-1)[:32].cpu())


def test_nan_and_inf_tokens_and_module_forward():
    from vq_gan_b200 import VectorQuantizer, ops
    z, E = _collapsed(4096, 64, 8, 1e-4, 16384, 0.05, seed=7)
    z[::97] = float("nan")
    z[5::1013] = float("inf")
    zc, Ec = _as_images(z, 1024).cuda(), E.cuda()
    idx, dmin, st = ops.search(zc, Ec, 4)
    ref_idx, ref_dmin, _ = ops.search(zc, Ec, 2)
    assert st.tolist()[3] > 0
    assert torch.equal(idx, ref_idx)
    assert torch.equal(dmin.nan_to_num(nan=-7.0), ref_dmin.nan_to_num(nan=-7.0))
    # through the module (automatic kernel choice), forward + backward, against the CPU oracle on clean tokens
    from oracle import vq_oracle as orc
    z, E = _collapsed(4096, 64, 8, 1e-4, 16384, 0.05, seed=8)
    zi = _as_images(z, 1024)
    vq = VectorQuantizer(4096, 64, 0.25).cuda()
    with torch.no_grad():
        vq.embedding.weight.copy_(E)
    zin = zi.cuda().requires_grad_(True)
    z_q, loss_dict, indices = vq(zin)
    loss_dict["vq_loss"].backward()
    assert vq.last_search_stats.tolist()[1] == 4 and vq.last_search_stats.tolist()[3] > 0
    rep = orc.compare_indices(indices.reshape(-1).cpu(), orc.search_with_gap(orc.tokens_of(zi), E))
    assert rep["outside"] == 0, rep
    fo = orc.forward(zi, E, 0.25, idx=indices.reshape(-1).cpu())
    assert torch.equal(z_q.detach().cpu(), fo["z_q"])


def test_cuda_graph_replay_with_the_tier_inside():
    """The tier decides everything on the device, so a captured forward replays correctly whether the data make it run
    (collapsed codebook), decline, or stay idle (healthy codebook) -- one graph, three codebooks."""
    from vq_gan_b200 import VectorQuantizer
    from vq_gan_b200.graphs import GraphedVectorQuantizer
    K, D, B, H = 4096, 64, 16, 32
    torch.manual_seed(3)
    vq = VectorQuantizer(K, D, 0.25, lazy_stats=True).cuda()
    gvq = GraphedVectorQuantizer(vq, torch.randn(B, D, H, H, device="cuda", requires_grad=True))
    z_c, E_c = _collapsed(K, D, 8, 1e-4, B * H * H, 0.05, seed=11)
    z_1, E_1 = _collapsed(K, D, 1, 1e-4, B * H * H, 0.05, seed=12)
    cases = ((z_c, E_c, "taken"), (z_1, E_1, "declined"), (torch.randn(B * H * H, D), torch.randn(K, D), "idle"))
    for z_rows, E, what in cases:
        with torch.no_grad():
            vq.embedding.weight.copy_(E)
        z = _as_images(z_rows, H * H).view(B, D, H, H).cuda()
        g = torch.randn(B, D, H, H, device="cuda")
        outs = []
        for mod in (vq, gvq):
            zc = z.clone().requires_grad_(True)
            vq.embedding.weight.grad = None
            z_q, ld, idx = mod(zc)
            torch.autograd.backward((z_q, ld["vq_loss"]), (g, torch.ones((), device="cuda")))
            outs.append((z_q.detach().clone(), idx.clone(), zc.grad.clone(), vq.last_search_stats.tolist()))
        e, r = outs
        assert torch.equal(e[0], r[0]) and torch.equal(e[1], r[1]) and torch.equal(e[2], r[2]), what
        st = e[3]
        assert st[1] == 4, st
        if what == "taken":
            assert st[3] == st[0] > 0, st
        elif what == "declined":
            assert st[3] == 0 and st[0] > 0.9 * B * H * H, st
        else:
            assert st[3] == 0 and st[0] < 0.01 * B * H * H, st
