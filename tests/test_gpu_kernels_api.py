"""Kernel-level API of the reference (kernels/fused_edge_features.py, kernels/fused_rbf_cutoff.py) on the
CUDA stage kernels: the reference's own self-tests (fused_edge_features.py:210-258, fused_rbf_cutoff.py:250-290)
restated with the reference's own tolerances, plus bit-level checks against the model's arithmetic, and
``radius_graph(positions, r, batch)`` called exactly like the reference calls it (student_model.py:165-171)."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden
from oracle import painn_oracle as po

pytestmark = pytest.mark.gpu


def _edge_features_pytorch(positions, edge_index, eps=1e-8):
    """kernels/fused_edge_features.py:169-199 (the reference's own baseline)."""
    src, dst = edge_index
    vec = positions[src] - positions[dst]
    dist = torch.norm(vec, dim=1)
    return vec, dist, vec / (dist.unsqueeze(1) + eps)


def _rbf_cutoff_pytorch(distances, centers, gamma, r_cut):
    """kernels/fused_rbf_cutoff.py:203-243."""
    rbf = torch.exp(-gamma * (distances.unsqueeze(-1) - centers.unsqueeze(0)) ** 2)
    cut = torch.where(distances < r_cut, 0.5 * (torch.cos(np.pi * distances / r_cut) + 1.0), torch.zeros_like(distances))
    return rbf * cut.unsqueeze(-1)


def test_fused_edge_features_reference_self_test():
    from mlff_distiller_b200.kernels import fused_edge_features_triton
    torch.manual_seed(42)
    for n_atoms, n_edges in ((12, 132), (5000, 100_003), (3, 0)):
        positions = torch.randn(n_atoms, 3, device="cuda")
        src = torch.randint(0, n_atoms, (n_edges,), device="cuda")
        dst = torch.randint(0, n_atoms, (n_edges,), device="cuda")
        edge_index = torch.stack([src, dst], dim=0)
        ref = _edge_features_pytorch(positions, edge_index)
        out = fused_edge_features_triton(positions, edge_index)
        for a, b in zip(ref, out):
            assert a.shape == b.shape
            assert torch.allclose(a, b, atol=1e-4, rtol=1e-3)          # the reference's tolerance (:241-249)
        # the model's own eps placement, to rounding: same ops as student_model.py:706-715
        out_m = fused_edge_features_triton(positions, edge_index, eps_placement="model")
        keep = ref[1] > 1e-3                                            # self pairs (src == dst) divide 0 / eps
        assert torch.equal(out_m[0], ref[0])
        assert torch.allclose(out_m[1][keep], ref[1][keep], rtol=3e-7, atol=0)
        assert torch.allclose(out_m[2][keep], ref[2][keep], rtol=0, atol=3e-7)


def test_fused_rbf_cutoff_reference_self_test_and_module():
    from mlff_distiller_b200.kernels import FusedRBFCutoff, fused_rbf_cutoff_triton
    torch.manual_seed(42)
    for n_edges, n_rbf, cutoff in ((132, 20, 5.0), (70_001, 12, 5.0), (257, 10, 4.0)):
        distances = torch.rand(n_edges, device="cuda") * cutoff * 1.1       # some beyond the cutoff
        distances[:2] = torch.tensor([cutoff, 0.0])
        centers = torch.linspace(0, cutoff, n_rbf, device="cuda")
        gamma = (1.0 / (torch.ones(n_rbf) * (cutoff / n_rbf))[0] ** 2).item()
        ref = _rbf_cutoff_pytorch(distances, centers, gamma, cutoff)
        out = fused_rbf_cutoff_triton(distances, centers, gamma, cutoff)
        assert out.shape == (n_edges, n_rbf)
        assert torch.allclose(ref, out, atol=1e-5, rtol=1e-4)              # the reference's tolerance (:283-285)
        assert float(out[distances >= cutoff].abs().max()) == 0.0          # strict d < r_cut
        mod = FusedRBFCutoff(num_rbf=n_rbf, cutoff=cutoff).cuda()
        assert set(dict(mod.named_buffers())) == {"centers", "widths"}
        assert torch.equal(mod(distances), out)


def test_stage_kernels_reproduce_the_oracles_edge_features():
    """The same two calls chained like StudentForceFieldOptimized.forward chains them
    (student_model_optimized.py:117-140) against the oracle's edge features on a golden structure."""
    from mlff_distiller_b200.kernels import fused_edge_features_triton, fused_rbf_cutoff_triton
    gold = load_golden("original")
    pos = torch.from_numpy(gold["drug50_positions"]).cuda()
    ei = torch.from_numpy(gold["drug50_edge_index"]).cuda()
    vec, dist, unit = fused_edge_features_triton(pos, ei, eps_placement="model")
    centers = torch.linspace(0, 5.0, 20)
    rbf = fused_rbf_cutoff_triton(dist, centers, 16.0, 5.0)
    w = {"rbf.centers": centers.double(), "rbf.widths": torch.full((20,), 0.25, dtype=torch.float64)}
    d_ref, u_ref, rbf_ref = po.edge_features(w, (pos[ei[0]] - pos[ei[1]]).double().cpu(), 5.0)
    assert float((dist.double().cpu() - d_ref).abs().max()) < 1e-6
    assert float((unit.double().cpu() - u_ref).abs().max()) < 1e-6
    assert float((rbf.double().cpu() - rbf_ref).abs().max()) < 1e-6


def test_radius_graph_with_the_reference_call_signature():
    """radius_graph(positions, r, batch) with no engine argument (student_model.py:165-171), CUDA positions."""
    from mlff_distiller_b200.student_model import radius_graph
    gold = load_golden("original")
    for case in golden_cases(gold):
        pos = torch.from_numpy(gold[f"{case}_positions"]).cuda()
        batch = po.batch_from_offsets(gold[f"{case}_offsets"]).cuda()
        ei = radius_graph(pos, 5.0, batch)
        assert ei.dtype == torch.int64 and ei.device.type == "cuda"
        assert np.array_equal(ei.cpu().numpy(), gold[f"{case}_edge_index"]), case
        ei2 = radius_graph(pos, r=5.0, batch=batch, loop=False, use_torch_cluster=False)
        assert torch.equal(ei, ei2)
    # a single structure without batch, a different cutoff, and self loops
    pos = torch.from_numpy(gold["benzene_positions"]).cuda()
    ei3 = radius_graph(pos, 3.0)
    ref3 = po.radius_graph_dense(pos.cpu(), 3.0).numpy()
    assert np.array_equal(ei3.cpu().numpy(), ref3)
    ei_loop = radius_graph(pos, 3.0, loop=True).cpu().numpy()
    assert ei_loop.shape[1] == ref3.shape[1] + 12
    order = np.lexsort((ei_loop[1], ei_loop[0]))
    assert np.array_equal(order, np.arange(ei_loop.shape[1]))
    with pytest.raises(RuntimeError, match="CUDA only"):
        radius_graph(pos.cpu(), 3.0)
