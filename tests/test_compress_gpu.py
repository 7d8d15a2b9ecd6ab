"""GPU: the pipelined host-buffer path (compress_groups_from_host) produces factors as good as the HBM-resident
path (compress_groups) and returns them in the caller's pinned host buffers.  (The factors are not bit-equal:
the Gaussian test matrix of a factorisation is seeded by the matrix's position in its batch, and the pipeline cuts
the groups into different batches.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chunk_groups", [1, 2, 3])
def test_host_pipeline_equals_resident_path(chunk_groups):
    from xkv_b200 import compress, synthetic

    dev = torch.device("cuda", 0)
    g, h, s, d, ng, rk, rv = 2, 2, 1024, 64, 3, 64, 96
    keys = [synthetic.make_group_kv(g, h, s, d, 1.0, 10 + i, device=dev) for i in range(ng)]
    vals = [synthetic.make_group_kv(g, h, s, d, 0.5, 20 + i, device=dev) for i in range(ng)]
    ref = compress.compress_groups(keys, vals, rk, rv)
    h_keys = [[t.transpose(1, 2).contiguous().cpu().pin_memory() for t in grp] for grp in keys]
    h_vals = [[t.transpose(1, 2).contiguous().cpu().pin_memory() for t in grp] for grp in vals]
    n = g * h * d
    host_out = []
    for _ in range(ng):
        for r in (rk, rv):
            host_out.append(torch.empty(s, r, dtype=torch.bfloat16).pin_memory())
            host_out.append(torch.empty(r, n, dtype=torch.bfloat16).pin_memory())
    out, host = compress.compress_groups_from_host(h_keys, h_vals, rk, rv, dev, chunk_groups=chunk_groups,
                                                   host_out=host_out)
    torch.cuda.synchronize()
    assert len(out) == ng and len(host) == 4 * ng
    k = 0
    for gi, (a, b) in enumerate(zip(out, ref)):
        for fa, fb, layers in ((a.key, b.key, keys[gi]), (a.value, b.value, vals[gi])):
            x = torch.cat(layers, dim=1).transpose(1, 2).reshape(s, n).double()
            ea = (torch.linalg.norm(x - fa.reconstruct().double()) / torch.linalg.norm(x)).item()
            eb = (torch.linalg.norm(x - fb.reconstruct().double()) / torch.linalg.norm(x)).item()
            assert abs(ea - eb) <= 2e-3 * eb, (ea, eb)
            assert torch.equal(host[k], fa.A.cpu()) and torch.equal(host[k + 1], fa.Vt.cpu())
            k += 2
