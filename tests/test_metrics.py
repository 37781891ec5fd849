"""Frame metrics: the oracle restatement against values produced by the reference's own
`psnr` / `ssim` (tests/golden/metrics.npz, oracle/gen_golden_metrics.py), and the CUDA path against both."""
import math

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M

GOLD = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "metrics.npz"))
CASES = ["a", "b", "c", "d", "same"]
SSIM_TOL = 1e-10   # fp64 both sides; summation order and (in the reference) OpenCV's DFT-based filter2D differ
PSNR_TOL = 1e-9    # dB


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_values(name):
    p, s = M.frame_metrics(GOLD[name + "_pred"].astype(np.float64), GOLD[name + "_gt"])
    assert abs(p - float(GOLD[name + "_psnr"])) < PSNR_TOL
    assert abs(s - float(GOLD[name + "_ssim"])) < SSIM_TOL


def test_oracle_edge_cases():
    a = np.full((12, 13, 3), 7.0)
    assert M.psnr(a, a) == float("inf")                      # utils.py:657-658
    assert abs(M.ssim(a, a) - 1.0) < 1e-12
    with pytest.raises(ValueError):
        M.ssim(a, a[:-1])
    g = M.gaussian_window_1d()
    assert g.shape == (11,) and abs(g.sum() - 1) < 1e-15 and np.allclose(g, g[::-1])
    # np.around is round-half-to-even: 0.5 -> 0, 1.5 -> 2, 2.5 -> 2
    x = (np.array([0.5, 1.5, 2.5]) / 255 * 2 - 1).reshape(3, 1, 1)
    q = M.quantise_prediction(np.broadcast_to(x, (3, 1, 1)))
    assert q.shape == (1, 1, 3)


def test_average_class_bookkeeping():
    from demfi_b200.metrics import AverageClass
    m = AverageClass("PSNR:", ":6.3f")
    m.update(30.0)
    m.update(40.0, 3)
    assert m.val == 40.0 and m.count == 4 and abs(m.avg - 37.5) < 1e-12
    assert str(m) == "PSNR: 40.000 (avg:37.500)"


def test_metrics_refuse_cpu_tensors():
    from demfi_b200 import metrics
    from demfi_b200._abi import DemfiError
    with pytest.raises(DemfiError):
        metrics.frame_metrics(torch.zeros(3, 16, 16), torch.zeros(3, 16, 16))
    with pytest.raises(ValueError):
        metrics.metric_sums(torch.zeros(3, 16, 16), torch.zeros(3, 16, 17))


# ----------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_values(name):
    from demfi_b200 import metrics
    dev = torch.device("cuda:0")
    pred = torch.from_numpy(GOLD[name + "_pred"]).to(dev)
    gt = torch.from_numpy(GOLD[name + "_gt"]).to(dev)
    p, s = metrics.frame_metrics(pred, gt)
    assert abs(p - float(GOLD[name + "_psnr"])) < PSNR_TOL
    assert abs(s - float(GOLD[name + "_ssim"])) < SSIM_TOL
    assert abs(metrics.psnr(gt, pred) - p) == 0 and abs(metrics.ssim(gt, pred) - s) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(11, 11), (11, 64), (27, 26), (37, 53), (64, 96), (720, 1280)])
def test_cuda_matches_oracle_shapes(h, w):
    """Ragged sizes around the 16-wide tiles and the minimum 11 x 11 image; 720p against the oracle (a few seconds of numpy)."""
    from demfi_b200 import metrics
    rng = np.random.default_rng(h * 1000 + w)
    gt = (rng.random((3, h, w)) * 2.4 - 1.2).astype(np.float32)
    pred = (gt + 0.05 * rng.standard_normal(gt.shape)).astype(np.float32)
    dev = torch.device("cuda:0")
    p, s = metrics.frame_metrics(torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev))
    po, so = M.frame_metrics(pred.astype(np.float64), gt)
    assert abs(p - po) < PSNR_TOL and abs(s - so) < SSIM_TOL


@pytest.mark.gpu
def test_cuda_batch_identity_and_repeatability():
    from demfi_b200 import metrics
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    gt = (torch.rand(4, 3, 40, 72, generator=g) * 2 - 1).to(dev)
    pred = (gt + 0.1 * torch.randn(4, 3, 40, 72, generator=g).to(dev)).contiguous()
    r = metrics.frame_metrics(pred, gt)
    for b in range(4):   # a batch is scored image by image
        assert metrics.frame_metrics(pred[b], gt[b]) == r[b]
    # fixed summation order: bitwise repeatable
    assert torch.equal(metrics.metric_sums(pred, gt), metrics.metric_sums(pred, gt))
    # a prediction compared with itself (both quantised the same way): zero error, SSIM 1, PSNR inf
    p, s = metrics.frame_metrics(pred[0], pred[0], target_is_prediction=True)
    assert p == float("inf") and abs(s - 1.0) < 1e-12
    # symmetry of both metrics when both sides are quantised alike
    a = metrics.frame_metrics(pred[0], pred[1], target_is_prediction=True)
    b = metrics.frame_metrics(pred[1], pred[0], target_is_prediction=True)
    assert abs(a[0] - b[0]) < 1e-12 and abs(a[1] - b[1]) < 1e-12


@pytest.mark.gpu
def test_score_forward_names_and_values():
    """main.py:757-838: twelve numbers per (pair, t), from the tensors the forward returned."""
    from demfi_b200 import metrics, synth
    from demfi_b200.DeMFInet import DeMFInet
    dev = torch.device("cuda:0")
    net = DeMFInet(synth.default_args()).to(dev).eval()
    net.load_state_dict(synth.make_state_dict(0))
    x = synth.make_frames(64, 96, 0).to(dev)
    with torch.no_grad():
        res = net(x, torch.tensor([[0.5]], device=dev), 2)
    gts = [x[:, :, 0].contiguous(), x[:, :, 1].contiguous(), res[4].contiguous()]   # stand-ins for S0/S1/St ground truths
    sc = metrics.score_forward(res, *gts)
    assert len(sc) == 12 and all(math.isfinite(v) for v in sc.values())
    want = M.frame_metrics(res[1][-1][2][0].double().cpu().numpy(), gts[2][0].cpu().numpy())
    assert abs(sc["intp_test_psnr"] - want[0]) < PSNR_TOL and abs(sc["intp_test_ssim"] - want[1]) < SSIM_TOL
    want = M.frame_metrics(res[0][0][0].double().cpu().numpy(), gts[0][0].cpu().numpy())
    assert abs(sc["test_psnr_S0_prime"] - want[0]) < PSNR_TOL and abs(sc["test_ssim_S0_prime"] - want[1]) < SSIM_TOL
