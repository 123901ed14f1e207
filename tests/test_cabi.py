"""The C-ABI library builds, loads, and exports every symbol include/ugl.h declares (no compute
calls here — those need a GPU); argument validation that runs before any launch is exercised too."""
import ctypes as C
import os
import re

import pytest

from unsupervised_depth_opticalflow_egomotion_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _cabi.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ugl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ugl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    syms = _declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
        assert s in _cabi.SIGNATURES, "no ctypes signature for " + s
    assert set(_cabi.SIGNATURES) == set(syms)


def test_version(lib):
    assert lib.ugl_version() == 100


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_struct_layout_matches_header(lib):
    # 3 ints + 2*6 ints, then pointer arrays: the C compiler pads to 8 before the first pointer
    a = _cabi.UglFlowLossArgs
    assert a.height.offset == 12 and a.width.offset == 36
    assert a.img_l.offset == 64 and a.img_l.size == 48
    assert C.sizeof(a) == 64 + 5 * 48 + 3 * 8 + 2 * 48 + 8 + 8 + 8 + 48


def test_struct_layouts_match_the_c_compiler(tmp_path):
    """sizeof / offsetof of every argument struct as gcc lays out include/ugl.h == the ctypes mirrors"""
    import os
    import subprocess
    structs = {"UglFlowLossArgs": _cabi.UglFlowLossArgs, "UglDepthPhotoArgs": _cabi.UglDepthPhotoArgs, "UglGeomFlowArgs": _cabi.UglGeomFlowArgs,
               "UglDispSmoothArgs": _cabi.UglDispSmoothArgs, "UglGeomRigidArgs": _cabi.UglGeomRigidArgs,
               "UglPyramidArgs": _cabi.UglPyramidArgs, "UglDepthSsimArgs": _cabi.UglDepthSsimArgs,
               "UglDepthPhotoGradArgs": _cabi.UglDepthPhotoGradArgs,
               "UglDepthConsisArgs": _cabi.UglDepthConsisArgs}
    lines = []
    for name, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for field in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, field[0], name, field[0]))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ugl.h"\nint main(void) {\n%s\nreturn 0; }\n' % "\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(__file__), "..", "include")
    subprocess.check_call(["gcc", "-I", inc, "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for field in cls._fields_:
            assert int(got["%s.%s" % (name, field[0])]) == getattr(cls, field[0]).offset, (name, field[0])


def test_argument_validation_happens_before_launch(lib):
    a = _cabi.UglFlowLossArgs()
    a.batch, a.levels, a.scales = 1, 1, 2            # scales > levels
    assert lib.ugl_flow_loss_forward(C.byref(a)) == -1
    assert b"scales" in lib.ugl_last_error()
    a.scales = 1
    a.height[0], a.width[0] = 2, 2                   # too small for the second-order stencil
    assert lib.ugl_flow_loss_forward(C.byref(a)) == -4
    a.height[0], a.width[0] = 8, 8                   # null input pointers
    assert lib.ugl_flow_loss_forward(C.byref(a)) == -1
    assert lib.ugl_image_pyramid(None, 1, 3, 8, 8, 2, 0, None, None) == -1
    assert lib.ugl_warp_flow_forward(None, None, 1, 3, 8, 8, 0, None, None, None) == -1


def test_fused_entries_validate_before_launch(lib):
    """every struct-taking entry rejects null / inconsistent arguments on the host (no GPU needed) with UGL_EINVAL and a message"""
    for name in ("ugl_geom_flow_forward_grad", "ugl_geom_flow_combine", "ugl_depth_photo_forward", "ugl_depth_photo_backward",
                 "ugl_disp_smooth_forward_grad", "ugl_disp_smooth_combine", "ugl_geom_rigid_forward", "ugl_geom_rigid_backward",
                 "ugl_image_pyramid_multi", "ugl_flow_loss_forward_grad", "ugl_flow_loss_combine"):
        assert getattr(lib, name)(None) == -1, name
        assert b"null" in lib.ugl_last_error(), name
    g = _cabi.UglGeomFlowArgs()
    g.flow.batch, g.flow.levels, g.flow.scales = 1, 3, 4
    assert lib.ugl_geom_flow_forward_grad(C.byref(g)) == -1 and b"scales" in lib.ugl_last_error()
    d = _cabi.UglDispSmoothArgs()
    d.batch, d.lists, d.levels, d.height, d.width = 1, 4, 1, 8, 8               # more lists than UGL_DISP_SMOOTH_MAX_LISTS
    assert lib.ugl_disp_smooth_forward_grad(C.byref(d)) == -1
    d.lists = 1
    d.lheight[0], d.lwidth[0] = 3, 8                                              # level does not divide the full size
    assert lib.ugl_disp_smooth_forward_grad(C.byref(d)) == -4
    p = _cabi.UglPyramidArgs()
    p.batch, p.channels, p.height, p.width, p.levels, p.images = 1, 3, 8, 8, 5, 1    # levels outside 2..4
    assert lib.ugl_image_pyramid_multi(C.byref(p)) == -4
    r = _cabi.UglGeomRigidArgs()
    r.batch, r.height, r.width = 0, 8, 8
    assert lib.ugl_geom_rigid_forward(C.byref(r)) == -1
    down = (C.c_float * 1)(1.0)
    assert lib.ugl_pose_setup_forward(None, None, None, down, 1, 2, 1, None, None, None, None) == -1
    assert lib.ugl_pose_setup_forward(C.c_void_p(16), C.c_void_p(16), None, down, 1, 3, 1, None, None, None, None) == -1   # n > 2
    assert lib.ugl_geom_rigid_workspace_bytes(2, 8, 8) == 2 * 1 * 42 * 4
    assert lib.ugl_disp_smooth_fused_workspace_bytes(C.byref(d)) == 1 * 1 * 1 * 2 * 4


def test_ops_refuse_cpu_tensors():
    import torch
    from unsupervised_depth_opticalflow_egomotion_b200 import ops
    x = torch.rand(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.warp_flow(x, torch.zeros(1, 2, 8, 8))
    with pytest.raises(ValueError, match="not equal to the shape of flow"):
        ops.warp_flow(x, torch.zeros(1, 2, 8, 9))
