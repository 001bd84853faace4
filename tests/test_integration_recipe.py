"""INTEGRATION.md section 1 against the real reference checkout (only where /root/reference exists: the authoring
container; the GPU box does not have it).  Import-level: no kernel is launched."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not present")


def test_dropin_install_keeps_the_rest_of_the_reference_importable():
    code = textwrap.dedent("""
        import sys, types, warnings
        warnings.filterwarnings("ignore")
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        m = types.ModuleType("matplotlib"); m.cm = types.ModuleType("matplotlib.cm")      # not installed here
        sys.modules["matplotlib"] = m; sys.modules["matplotlib.cm"] = m.cm
        import rrnet_b200.host.dropin as dropin
        done = dropin.install()
        assert len(done) == len(dropin.WHOLE) + len(dropin.SYMBOLS)
        import rrnet_b200.host as H
        # the reference's own names now resolve to the mirror
        import models.rrnet, operators.rrnet_operator, ext.nms.nms_wrapper
        assert models.rrnet.RRNet.__module__.startswith("rrnet_b200.host")
        assert operators.rrnet_operator.RRNetOperator.__module__.startswith("rrnet_b200.host")
        assert ext.nms.nms_wrapper.soft_nms.__module__.startswith("rrnet_b200.host")
        # patched symbols, and the untouched neighbours in the same reference modules
        import modules.loss.focalloss as FL, modules.loss.functional as LF, modules.loss.regl1loss as RL
        assert FL.FocalLossHM.__module__.startswith("rrnet_b200.host") and FL.FocalLoss.__module__ == "modules.loss.focalloss"
        assert LF.focal_loss_for_hm.__module__.startswith("rrnet_b200.host")
        for keep in ("focal_loss", "giou_loss", "kl_loss", "flat_tensor"):
            assert getattr(LF, keep).__module__ == "modules.loss.functional", keep
        assert RL.RegL1Loss.__module__.startswith("rrnet_b200.host")
        import datasets.transforms.functional as TF, datasets.transforms.transforms as TT
        # the per-sample render runs in forked DataLoader workers: it stays the reference's CPU code by default
        assert TF.to_heatmap.__module__ == "datasets.transforms.functional" and callable(TF.denormalize)
        assert TT.ToHeatmap.__module__ == "datasets.transforms.transforms"
        done2 = dropin.install(gpu_targets=True)
        assert len(done2) == len(done) + len(dropin.GPU_TARGET_SYMBOLS)
        assert TT.ToHeatmap.__name__ == "DeferredToHeatmap" and TT.ToHeatmap.__module__.startswith("rrnet_b200.host")
        import torch
        out = TT.ToHeatmap()((torch.zeros(3, 64, 64), torch.ones(5, 8)))          # CPU only, no CUDA in the worker
        assert out[2].numel() == 0 and out[6].shape == (5, 1) and not any(t.is_cuda for t in out)
        import utils.metrics.metrics as UM
        assert UM.get_tp.__module__.startswith("rrnet_b200.host") and UM.calculate_ap_rc.__module__ == "utils.metrics.metrics"
        # an operator of the reference that is NOT on the path still imports, and gets the mirror's losses / NMS
        import operators.centernet_operator as CO
        assert CO.FocalLossHM is FL.FocalLossHM and CO.RegL1Loss is RL.RegL1Loss
        assert CO.soft_nms is ext.nms.nms_wrapper.soft_nms
        print("recipe ok")
    """) % (REF, REPO)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "recipe ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
