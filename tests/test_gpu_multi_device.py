"""Tensors on a device that is NOT torch.cuda.current_device(): the reference's train_parsenet_e2e.py keeps the fit stage
on cuda:alt_gpu (= 1 on a multi-GPU box, :58) while the process's current device stays cuda:0.  Every C-ABI launch must
then run on the device (and torch stream) that owns its pointers.  Needs 2 GPUs (`gpurun --gpus 2`), skipped otherwise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

needs2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")


@needs2
def test_fitting_loss_on_cuda1_while_current_device_is_cuda0():
    from oracle.make_golden_helpers import e2e_inputs
    from src.residual_utils import Evaluation
    from test_gpu_fitting import _seeded_splinenet
    N = 2000
    pts, nrm, lab, prim, emb, logp = e2e_inputs(N, 77, False)
    out = {}
    for dev_i in (0, 1):
        torch.cuda.set_device(0)                       # the current device never changes
        dev = torch.device("cuda", dev_i)
        ev = Evaluation(open_decoder=_seeded_splinenet(0, 41), closed_decoder=_seeded_splinenet(1, 42))   # on cuda:0
        E = emb.to(dev).requires_grad_()
        np.random.seed(5)
        res, extra = ev.fitting_loss(E, torch.from_numpy(pts).to(dev), torch.from_numpy(nrm).to(dev), lab, prim.copy(),
                                     logp.to(dev), quantile=0.015, iterations=10, lamb=0.1)
        assert res[0].device == dev
        assert next(ev.fitter.open_control_decoder.parameters()).device == dev     # decoders followed the data
        res[0].backward()
        torch.cuda.synchronize(dev)
        out[dev_i] = (float(res[0]), E.grad.cpu().clone(), extra[1].copy())
        assert torch.cuda.current_device() == 0
    # same kernels, same inputs; floating-point atomics (moment / statistics accumulation) order differently per run
    assert abs(out[0][0] - out[1][0]) <= 1e-5 * abs(out[0][0]), (out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][2], out[1][2])
    scale = out[0][1].abs().max().item()
    assert (out[0][1] - out[1][1]).abs().max().item() <= 1e-4 * scale


@needs2
def test_segnet_on_cuda1_while_current_device_is_cuda0():
    from src.PointNet import PrimitivesEmbeddingDGCNGn
    from src.segment_loss import EmbeddingLoss, primitive_loss
    from tools.synth import synth_cloud
    B, N, k = 2, 512, 16
    pts, nrm, lab, prim = synth_cloud(B, N, seed=3)
    x = torch.from_numpy(np.concatenate([pts, nrm], 2)).permute(0, 2, 1).contiguous()
    got = {}
    for dev_i in (0, 1):
        torch.cuda.set_device(0)
        dev = torch.device("cuda", dev_i)
        torch.manual_seed(0)
        m = PrimitivesEmbeddingDGCNGn(embedding=True, emb_size=128, primitives=True, num_primitives=10,
                                      loss_function=EmbeddingLoss(margin=1.0).triplet_loss, mode=5, num_channels=6,
                                      nn_nb=k).to(dev)
        np.random.seed(0)
        emb, lp, el = m(x.to(dev), torch.from_numpy(lab).to(dev), True)
        (el.sum() + primitive_loss(lp, torch.from_numpy(prim).to(dev))).backward()
        torch.cuda.synchronize(dev)
        got[dev_i] = (emb.detach().cpu(), m.conv1.weight.grad.cpu())
        assert torch.cuda.current_device() == 0
    torch.testing.assert_close(got[0][0], got[1][0], rtol=0, atol=0)
    torch.testing.assert_close(got[0][1], got[1][1], rtol=1e-5, atol=1e-7)     # (fp64 atomics in the statistics)


@needs2
def test_mixed_device_arguments_raise():
    from pnb200 import cabi, ops
    a = torch.randn(1, 256, 6, device="cuda:0")
    W = torch.randn(64, 6, device="cuda:1")
    with pytest.raises(cabi.PnError, match="different devices"):
        ops.linear_fwd(a, W)
