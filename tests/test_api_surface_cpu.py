"""CPU: the reference-facing error behaviour and the pure-host helpers of the drop-in classes (SURVEY.md 8b "Errors" row).
Nothing here launches a kernel; what needs the GPU is in test_gpu_*.py."""
import numpy as np
import pytest
import torch

import pdc_b200
from pdc_b200 import loss_composer
from pdc_b200.dense_correspondence_network import DenseCorrespondenceNetwork as DCN


def test_get_fcn_errors_like_the_reference():
    # dense_correspondence_network.py:360-383: unknown model class -> ValueError; only Resnet34_8s exists in this build
    with pytest.raises(ValueError):
        DCN.get_fcn({"backbone": {"model_class": "Transformer"}, "descriptor_dimension": 3})
    with pytest.raises(ValueError):
        DCN.get_fcn({"backbone": {"model_class": "Resnet", "resnet_name": "Resnet101_8s"}, "descriptor_dimension": 3})
    with pytest.raises(NotImplementedError):
        DCN.get_fcn({"backbone": {"model_class": "Unet"}, "descriptor_dimension": 3})


def test_from_config_needs_a_param_file_when_asked_to_load():
    # dense_correspondence_network.py:427: `assert model_param_file is not None`
    with pytest.raises(AssertionError):
        DCN.from_config({"descriptor_dimension": 3, "image_width": 64, "image_height": 64}, load_stored_params=True)


def test_get_loss_rejects_unknown_pair_type():
    # loss_composer.py:67
    pcl = pdc_b200.PixelwiseContrastiveLoss(image_shape=[8, 8], config={})
    e = loss_composer.empty_tensor()
    with pytest.raises(ValueError):
        loss_composer.get_loss(pcl, torch.tensor([7]), None, None, e, e, e, e, e, e, e, e)


def test_sentinel_helpers():
    # dense_correspondence_dataset_masked.py:209-223
    e = loss_composer.empty_tensor()
    assert e.dtype == torch.int64 and e.tolist() == [-1]
    assert loss_composer.is_empty(e) and not loss_composer.is_empty(torch.tensor([3])) and not loss_composer.is_empty(torch.tensor([-1, -1]))


def test_process_network_output_is_the_reference_view():
    # dense_correspondence_network.py:303-319: [N,D,H,W] -> [N, W*H, D] where (b,p,c) aliases pred[b,c,p // W, p % W]
    H, W, D, B = 4, 6, 3, 2
    dcn = DCN(fcn=torch.nn.Identity(), descriptor_dimension=D, image_width=W, image_height=H)
    pred = torch.arange(B * D * H * W, dtype=torch.float32).view(B, D, H, W)
    out = dcn.process_network_output(pred, B)
    assert out.shape == (B, W * H, D) and out.data_ptr() == pred.data_ptr()          # a view, no copy
    for (b, p, c) in [(0, 0, 0), (1, 7, 2), (0, 23, 1), (1, 13, 0)]:
        assert out[b, p, c] == pred[b, c, p // W, p % W]
    assert dcn.image_shape == [H, W] and dcn.descriptor_dimension == D


def test_find_best_match_is_numpy_argmin_with_first_minimum():
    # dense_correspondence_network.py:488-525
    rng = np.random.RandomState(0)
    H, W, D = 5, 7, 3
    res_a = rng.randn(H, W, D).astype(np.float32); res_b = rng.randn(H, W, D).astype(np.float32)
    res_b[3, 2] = res_a[1, 4]; res_b[4, 6] = res_a[1, 4]                            # two exact matches: the first (row-major) wins
    uv, diff, nd = DCN.find_best_match((4, 1), res_a, res_b)
    assert uv == (2, 3) and diff == 0.0 and nd.shape == (H, W)
    brute = np.sqrt(((res_b - res_a[1, 4]) ** 2).sum(axis=2))
    np.testing.assert_allclose(nd, brute, rtol=1e-6)
    uv2, diff2, _ = DCN.find_best_match_for_descriptor(res_a[0, 0], res_b)
    iy, ix = np.unravel_index(np.argmin(np.sqrt(((res_b - res_a[0, 0]) ** 2).sum(axis=2))), (H, W))
    assert uv2 == (ix, iy)


def test_clip_pixel_and_precision_switch():
    dcn = DCN(fcn=torch.nn.Identity(), descriptor_dimension=3, image_width=640, image_height=480)
    assert dcn.clip_pixel_to_image_size_and_round((639.6, 479.7)) == [639, 479]
    assert dcn.clip_pixel_to_image_size_and_round((10.4, 20.6)) == [10, 21]
    with pytest.raises(Exception):
        pdc_b200.set_default_precision("fp64")


def test_fused_adam_state_dict_is_torch_adam_compatible():
    """training.py:509-511 saves `optimizer.state_dict()` as NNNNNN.pth.opt and :147-150 resumes from it: a file written by
    torch.optim.Adam must load into FusedAdam (moments land at the right offsets of the flat arrays) and the other way round."""
    torch.manual_seed(0)
    net = pdc_b200.Resnet34_8s(num_classes=3)                      # CPU: only state handling is exercised, no kernel
    params = list(net.parameters())
    ref = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-4)
    for _ in range(2):
        for q in params:
            q.grad = torch.randn_like(q) * 1e-3
        ref.step()
    sd = ref.state_dict()
    opt = pdc_b200.FusedAdam(net, lr=1.0, weight_decay=0.0)
    opt.load_state_dict(sd)
    assert opt.step_count == 2 and opt.param_groups[0]["lr"] == 1e-4 and opt.param_groups[0]["weight_decay"] == 1e-4
    flat = net.flat_parameters
    for i, q in enumerate(params):
        off = (q.data_ptr() - flat.data_ptr()) // 4
        assert torch.equal(opt.exp_avg[off:off + q.numel()].view(q.shape), sd["state"][i]["exp_avg"])
        assert torch.equal(opt.exp_avg_sq[off:off + q.numel()].view(q.shape), sd["state"][i]["exp_avg_sq"])
    # alignment padding between tensors carries no state
    covered = sum(q.numel() for q in params)
    assert float(opt.exp_avg.abs().sum()) == pytest.approx(float(sum(sd["state"][i]["exp_avg"].abs().sum() for i in range(len(params)))), rel=1e-5)
    assert flat.numel() >= covered
    # and back: our checkpoint resumes a stock torch.optim.Adam
    back = torch.optim.Adam(params, lr=1.0)
    back.load_state_dict(opt.state_dict())
    sb = back.state_dict()
    assert sb["param_groups"][0]["lr"] == 1e-4 and len(sb["state"]) == len(params)
    for i in range(len(params)):
        assert torch.equal(sb["state"][i]["exp_avg"], sd["state"][i]["exp_avg"]) and float(sb["state"][i]["step"]) == 2.0
    # a checkpoint for another network is refused
    bad = {"state": {}, "param_groups": [dict(sd["param_groups"][0], params=[0, 1, 2])]}
    with pytest.raises(ValueError):
        opt.load_state_dict(bad)


def test_adjust_learning_rate_matches_the_reference_schedule():
    """training.py:544-558 with training.yaml's steps_between_learning_rate_decay / learning_rate_decay."""
    net = pdc_b200.Resnet34_8s(num_classes=3)
    opt = pdc_b200.FusedAdam(net, lr=1e-4, weight_decay=1e-4)
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1e-4)
    for it in range(1, 1001):
        pdc_b200.adjust_learning_rate(opt, it, 250, 0.9)
        if it % 250 == 0:                                           # the reference's body, verbatim semantics
            for g in ref.param_groups:
                g["lr"] = g["lr"] * 0.9
    assert abs(opt.param_groups[0]["lr"] - ref.param_groups[0]["lr"]) < 1e-18 and abs(opt.param_groups[0]["lr"] - 1e-4 * 0.9 ** 4) < 1e-12
