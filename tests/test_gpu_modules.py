"""GPU parity of the stand-alone PointNet++ modules (PointNetSetAbstraction,
PointNetSetAbstractionMsg, PointNetFeaturePropagation, index_points autograd) against the CPU
oracle's restatement of pointnet_util.py:166-320 on the same seeded inputs and the same checkpoint
tensors.  Tolerances (fp32 MLPs): outputs rtol 1e-3 / atol 1e-4 of the output scale; feature
gradients relative Frobenius error < 5e-3 (max-pool arg-max and ReLU routing make
the gradient piecewise: the oracle's own fp32 vs fp64 gradient differs by 5.9e-3)."""
import numpy as np
import pytest
import torch

from pointsecguard_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _load(module, sd, prefix):
    sub = {k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")}
    module.load_state_dict(sub)
    return module.cuda().eval()


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("arch,level", [("ssg", 1), ("ssg", 2), ("msg", 1), ("msg", 2)])
def test_set_abstraction_vs_oracle(arch, level):
    from oracle import pointnet2_oracle as PO
    from pointsecguard_b200.models import pointnet_util as PU
    sd = syn.make_state_dict(arch)
    npoint, radii, ks, mlps = syn.ARCH[arch]["sa"][level - 1]
    D = syn.ARCH[arch]["sa_in"][level - 1]
    B, N = 2, 2048 if level == 1 else 512
    g = torch.Generator().manual_seed(level)
    xyz = syn.make_blocks(B, N, level, "uniform")[:, :3].contiguous()
    pts = torch.randn(B, D, N, generator=g)
    if arch == "ssg":
        mod = PU.PointNetSetAbstraction(npoint if N > npoint else N // 4, radii[0], ks[0], D + 3, mlps[0], False)
    else:
        mod = PU.PointNetSetAbstractionMsg(npoint if N > npoint else N // 4, radii, ks, D, mlps)
    mod = _load(mod, sd, f"sa{level}")
    S = mod.npoint
    pc = pts.clone().requires_grad_(True)
    torch.manual_seed(11)
    ref_xyz, ref_out = PO.set_abstraction(sd, f"sa{level}", arch, S, radii, ks, [3] * len(radii), xyz, pc)
    up = torch.randn(ref_out.shape, generator=g)
    ref_out.backward(up)
    pg = pts.cuda().requires_grad_(True)
    torch.manual_seed(11)
    new_xyz, out = mod(xyz.cuda(), pg)
    assert torch.equal(new_xyz.cpu(), ref_xyz)                         # sampled coordinates bit-exact
    scale = ref_out.abs().max().item()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref_out.detach().numpy(), rtol=1e-3, atol=1e-4 * scale)
    out.backward(up.cuda())
    rel = _rel(pg.grad.cpu().numpy(), pc.grad.numpy())
    print(arch, level, "d points rel", rel)
    assert rel < 5e-3


@pytest.mark.parametrize("name,nl,D1,D2,N,S", [("fp1", 3, 0, 128, 1024, 256), ("fp2", 2, 64, 256, 512, 128),
                                              ("fp4", 2, 256, 512, 64, 16)])
def test_feature_propagation_vs_oracle(name, nl, D1, D2, N, S):
    from oracle import pointnet2_oracle as PO
    from pointsecguard_b200.models import pointnet_util as PU
    sd = syn.make_state_dict("ssg")
    B = 2
    g = torch.Generator().manual_seed(N)
    xyz1 = syn.make_blocks(B, N, 5, "uniform")[:, :3].contiguous()
    xyz2 = xyz1[:, :, torch.randperm(N, generator=g)[:S]].contiguous()   # coarse points are a subset (d2 = 0 hits)
    p1 = torch.randn(B, D1, N, generator=g) if D1 else None
    p2 = torch.randn(B, D2, S, generator=g)
    cin = D1 + D2
    mlp = [sd[f"{name}.mlp_convs.{j}.weight"].shape[0] for j in range(nl)]
    assert sd[f"{name}.mlp_convs.0.weight"].shape[1] == cin
    mod = _load(PU.PointNetFeaturePropagation(cin, mlp), sd, name)
    p1c = p1.clone().requires_grad_(True) if D1 else None
    p2c = p2.clone().requires_grad_(True)
    ref = PO.feature_propagation(sd, name, nl, xyz1, xyz2, p1c, p2c)
    up = torch.randn(ref.shape, generator=g)
    ref.backward(up)
    p1g = p1.cuda().requires_grad_(True) if D1 else None
    p2g = p2.cuda().requires_grad_(True)
    out = mod(xyz1.cuda(), xyz2.cuda(), p1g, p2g)
    scale = ref.abs().max().item()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-3, atol=1e-4 * scale)
    out.backward(up.cuda())
    r2 = _rel(p2g.grad.cpu().numpy(), p2c.grad.numpy())
    print(name, "d points2 rel", r2)
    assert r2 < 5e-3
    if D1:
        r1 = _rel(p1g.grad.cpu().numpy(), p1c.grad.numpy())
        print(name, "d points1 rel", r1)
        assert r1 < 5e-3


def test_feature_propagation_odd_widths_and_single_coarse_point():
    """Skip widths off the 16-column grid take the materialised-concat path; S == 1 repeats the
    single coarse point (pointnet_util.py:298-299)."""
    from pointsecguard_b200.models import pointnet_util as PU
    torch.manual_seed(3)
    mod = PU.PointNetFeaturePropagation(9 + 20, [32, 16])
    for bn in mod.mlp_bns:
        bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
    ref_mod = mod.eval()
    B, N, S = 2, 300, 40
    xyz1, xyz2 = torch.rand(B, 3, N), torch.rand(B, 3, S)
    p1, p2 = torch.randn(B, 9, N), torch.randn(B, 20, S)

    def ref_forward(x1, x2, a, b):
        from oracle import pointnet2_oracle as PO
        sd = {f"fp.{k}": v for k, v in ref_mod.state_dict().items()}
        return PO.feature_propagation(sd, "fp", 2, x1, x2, a, b)

    ref = ref_forward(xyz1, xyz2, p1, p2)
    out = mod.cuda()(xyz1.cuda(), xyz2.cuda(), p1.cuda(), p2.cuda())
    np.testing.assert_allclose(out.cpu().detach().numpy(), ref.detach().numpy(), rtol=1e-3, atol=1e-4)
    one = mod(xyz1.cuda(), xyz2[:, :, :1].cuda(), p1.cuda(), p2[:, :, :1].cuda())
    rep = p2[:, :, :1].repeat(1, 1, N)
    x = torch.cat([p1, rep], 1)
    m_cpu = PU.PointNetFeaturePropagation(9 + 20, [32, 16])
    m_cpu.load_state_dict({k: v.cpu() for k, v in mod.state_dict().items()})
    m_cpu.eval()
    for conv, bn in zip(m_cpu.mlp_convs, m_cpu.mlp_bns):
        x = torch.relu(bn(conv(x)))
    np.testing.assert_allclose(one.cpu().detach().numpy(), x.detach().numpy(), rtol=1e-3, atol=1e-4)


def test_index_points_backward_is_deterministic_scatter():
    from pointsecguard_b200.models import pointnet_util as PU
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(2, 500, 7, generator=g)
    idx = torch.randint(0, 500, (2, 64, 8), generator=g)
    up = torch.randn(2, 64, 8, 7, generator=g)
    pc = pts.clone().requires_grad_(True)
    pc[torch.arange(2).view(2, 1, 1), idx].backward(up)
    grads = []
    for _ in range(2):
        pg = pts.cuda().requires_grad_(True)
        PU.index_points(pg, idx.cuda()).backward(up.cuda())
        grads.append(pg.grad.clone())
    assert torch.equal(grads[0], grads[1])
    np.testing.assert_allclose(grads[0].cpu().numpy(), pc.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_sample_and_group_matches_oracle_layout():
    from oracle import pointnet2_oracle as PO
    from pointsecguard_b200.models import pointnet_util as PU
    x = syn.make_blocks(2, 1024, 9, "uniform")
    xyz = x[:, :3].permute(0, 2, 1).contiguous()
    pts = x.permute(0, 2, 1).contiguous()
    torch.manual_seed(2)
    new_xyz, new_points, grouped_xyz, fps_idx = PU.sample_and_group(256, 0.2, 32, xyz.cuda(), pts.cuda(), returnfps=True)
    torch.manual_seed(2)
    rf = PO.farthest_point_sample(xyz, 256)
    rn = PO.index_points(xyz, rf)
    ri = PO.query_ball_point(0.2, 32, xyz, rn)
    rg = PO.index_points(xyz, ri)
    ref_points = torch.cat([rg - rn.view(2, 256, 1, 3), PO.index_points(pts, ri)], -1)
    assert torch.equal(fps_idx.cpu(), rf) and torch.equal(new_xyz.cpu(), rn) and torch.equal(grouped_xyz.cpu(), rg)
    assert torch.equal(new_points.cpu(), ref_points)
