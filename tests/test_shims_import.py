"""The stand-ins make the reference's own modules importable.  Needs the reference tree
(this container only; skipped on the GPU box, where /root/reference does not exist)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scene")), reason="reference tree not present")
def test_reference_scene_modules_import_with_the_stand_ins():
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import seganygaussians_b200 as S; S.activate()\n"
        "import plyfile, simple_knn._C, pytorch3d.ops\n"
        "assert 'seganygaussians_b200' in plyfile.__file__ and 'seganygaussians_b200' in pytorch3d.ops.__file__\n"
        "import scene.gaussian_model, scene.gaussian_model_ff, scene.dataset_readers\n"
        "import diff_gaussian_rasterization_contrastive_f as cf, gaussian_renderer\n"
        "assert 'seganygaussians_b200' in cf.__file__ and 'seganygaussians_b200' in gaussian_renderer.__file__\n"
        "print('ok')\n" % (REF, ROOT))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
