"""CPU check of the ALGORITHM behind csrc/loss_fused.cu (tests/fused_model.py: tiling, base-2 statistics, bf16 rounding points,
bf16 dE partial tiles, the <dWs, What> identity) against the fp64 oracle, at the tolerances the GPU tests apply to the kernel.
Pins the precision contract of the bf16 loss path without a GPU."""
import pytest
import torch

from oracle import textreid_oracle as O
from textreid_b200.synthetic import loss_inputs
from tests.fused_model import fused_loss_model

ARGS = ("v_embed", "t_embed", "v_key", "t_key", "labels", "v_queue", "t_queue", "id_queue", "projection")
KEYS = ("instance_loss", "infonce_loss", "global_align_loss")


@pytest.mark.parametrize("N,D,K,C,masked", [(128, 256, 2048, 11003, "some"), (32, 64, 128, 1000, "empty"), (100, 192, 200, 257, "some"),
                                              (8, 128, 384, 130, "some")])
def test_fused_algorithm_meets_the_bf16_contract(N, D, K, C, masked):
    inp = loss_inputs(N, D, K, C, seed=N + K, masked=masked)
    losses, dv, dt, dw = fused_loss_model(*[inp[k] for k in ARGS], epsilon=0.1)
    ref_l, rv, rt, rp = O.moco_loss_dict_with_grads(*[inp[k].double() if inp[k].dtype.is_floating_point else inp[k] for k in ARGS],
                                                    epsilon=0.1)
    for k in KEYS:                                           # north star: 1e-3 on the bf16 path
        torch.testing.assert_close(losses[k].double(), ref_l[k], rtol=1e-3, atol=1e-4)
    for got, ref in ((dv, rv), (dt, rt), (dw, rp)):
        err = (got.double() - ref).abs().max() / ref.abs().max()
        assert float(err) < (2e-2 if N >= 32 else 4e-2), float(err)
        cos = torch.nn.functional.cosine_similarity(got.double().flatten(), ref.flatten(), dim=0)
        assert float(cos) > 0.9995, float(cos)


def test_fused_algorithm_all_slots_masked_and_no_smoothing():
    """K' = 0: only the positive logit is left, InfoNCE is exactly 0 and its gradient vanishes (SURVEY 8c F-edge)."""
    inp = loss_inputs(16, 64, 128, 300, seed=5)
    inp["id_queue"][:] = inp["labels"][0]
    losses, dv, dt, dw = fused_loss_model(*[inp[k] for k in ARGS], epsilon=0.0)
    ref_l, rv, rt, rp = O.moco_loss_dict_with_grads(*[inp[k].double() if inp[k].dtype.is_floating_point else inp[k] for k in ARGS],
                                                    epsilon=0.0)
    assert float(losses["infonce_loss"]) == 0.0 and float(ref_l["infonce_loss"]) == 0.0
    for k in KEYS:
        torch.testing.assert_close(losses[k].double(), ref_l[k], rtol=1e-3, atol=1e-4)
    for got, ref in ((dv, rv), (dt, rt), (dw, rp)):
        assert float((got.double() - ref).abs().max() / ref.abs().max()) < 4e-2


def test_column_dot_identity_used_by_the_dw_epilogue():
    """<E^T dz, What>_col == sum_rows dz * (E What): the identity that lets the kernel form the column-normalisation Jacobian
    without a second pass over W (exact in exact arithmetic; checked in fp64)."""
    g = torch.Generator().manual_seed(0)
    E = torch.randn(24, 16, generator=g, dtype=torch.float64)
    W = torch.randn(16, 40, generator=g, dtype=torch.float64)
    What = W / W.norm(dim=0)
    dz = torch.randn(24, 40, generator=g, dtype=torch.float64)
    lhs = ((E.t() @ dz) * What).sum(0)
    rhs = (dz * (E @ What)).sum(0)
    torch.testing.assert_close(lhs, rhs, rtol=1e-12, atol=1e-12)
