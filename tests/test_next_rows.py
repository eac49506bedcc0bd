"""Host-side pieces of the 'next' rows: compact index storage (N3) and sync-free stats (N4)."""
import io

import pytest
import torch

from vq_gan_b200 import indexio
from vq_gan_b200.stats import DeviceStats


@pytest.mark.parametrize("K,dt,nbytes", [(128, torch.uint8, 1), (256, torch.uint8, 1), (257, torch.uint16, 2),
                                         (16384, torch.uint16, 2), (65536, torch.uint16, 2),
                                         (65537, torch.int32, 4)])
def test_index_roundtrip(K, dt, nbytes):
    idx = torch.randint(0, K, (3, 7, 5), generator=torch.Generator().manual_seed(K))
    idx[0, 0, 0], idx[0, 0, 1] = 0, K - 1
    blob = indexio.pack_indices(idx, K)
    assert blob["codes"].dtype == dt and indexio.bytes_per_token(K) == nbytes
    buf = io.BytesIO()
    torch.save(blob, buf)
    buf.seek(0)
    back = indexio.unpack_indices(torch.load(buf))
    assert back.dtype == torch.int64 and torch.equal(back, idx)


def test_index_pack_rejects_bad_input():
    with pytest.raises(ValueError):
        indexio.pack_indices(torch.tensor([[[5]]]), 5)
    with pytest.raises(TypeError):
        indexio.pack_indices(torch.tensor([[[1]]], dtype=torch.int32), 5)


def test_device_stats_accumulate_and_flush():
    st = DeviceStats(8, "cpu")
    for i in range(4):
        ld = {"vq_loss": torch.tensor(1.0 + i), "codebook_loss": torch.tensor(0.5 * i)}
        st.update(ld, torch.tensor([1, 0, 0, 2, 0, 0, 0, 0]))
    out = st.flush()
    assert out["steps"] == 4 and out["tokens"] == 12
    assert abs(out["vq_loss"] - 2.5) < 1e-12 and abs(out["codebook_loss"] - 0.75) < 1e-12
    assert out["codebook_usage_ratio"] == 0.25
    assert st.flush()["steps"] == 0
