"""GPU: the append/project kernel against torch (x_new @ V with fp32 accumulation, rounded to bf16), and the
property that a projected token reconstructs as well as the prefill tokens do."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,n,r", [(1, 4096, 512), (1, 4096, 768), (3, 1024, 128), (11, 512, 64), (1, 200, 34)])
def test_append_project_matches_torch(T, n, r):
    from xkv_b200 import ops

    torch.manual_seed(T + n + r)
    x = torch.randn(T, n, device="cuda").bfloat16()
    v = (torch.randn(n, r, device="cuda") / n ** 0.5).bfloat16()
    got = ops.append_project(x, v)
    torch.cuda.synchronize()
    ref = x.float() @ v.float()
    assert got.shape == (T, r) and got.dtype == torch.bfloat16
    assert (got.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("T", [1, 8, 11])
def test_append_project_many_is_one_launch_and_matches_the_single_calls(T):
    """A group's K and V projection in one launch (config 2's shapes): bit-identical to two single calls."""
    from xkv_b200 import ops

    torch.manual_seed(T)
    n = 4096
    xs = [torch.randn(T, n, device="cuda").bfloat16() for _ in range(2)]
    vs = [(torch.randn(n, r, device="cuda") / n ** 0.5).bfloat16() for r in (512, 768)]
    outs = [torch.empty(T, r, dtype=torch.bfloat16, device="cuda") for r in (512, 768)]
    before = ops.launch_count()
    ops.append_project_many(xs, vs, outs)
    assert ops.launch_count() - before == (T + 7) // 8
    singles = [ops.append_project(x, v) for x, v in zip(xs, vs)]
    torch.cuda.synchronize()
    for o, s1, x, v in zip(outs, singles, xs, vs):
        assert torch.equal(o, s1)
        ref = x.float() @ v.float()
        assert (o.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


def test_projected_token_reconstructs_like_prefill_tokens():
    from xkv_b200 import factorize, ops, synthetic

    S, n, r = 2048, 1024, 128
    x = synthetic.group_matrix(S + 64, n, 1.0, seed=9, device="cuda")
    (f,) = factorize.factorize_batch([x[:S].contiguous()], r)
    a_new = ops.append_project(x[S:].contiguous(), f.V)
    torch.cuda.synchronize()
    rec_new = a_new.float() @ f.Vt.float()
    err_new = ((x[S:].float() - rec_new).norm() / x[S:].float().norm()).item()
    err_old = ((x[:S].float() - f.reconstruct().float()).norm() / x[:S].float().norm()).item()
    print(f"reconstruction error: prefill rows {err_old:.4f}, appended rows {err_new:.4f}")
    assert err_new < 1.5 * err_old     # same row space (shared right factor): appended tokens compress alike
