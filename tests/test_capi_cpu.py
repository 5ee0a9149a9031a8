"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/pixel_engine.h
declares, its ctypes binding covers them all, and it fails loudly without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

import pe_testlib as T

lb = pytest.importorskip("lives_b200")
from lives_b200 import _capi  # noqa: E402

HEADER = os.path.join(T.REPO, "include", "pixel_engine.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(pe_[a-z0-9_]+)\s*\(", src))
    names -= {n for n in names if n.endswith("_f")}  # function-pointer typedefs
    return names


def test_header_symbols_exported_and_bound():
    if not os.path.exists(_capi.LIB_PATH):
        from lives_b200.build import build
        build()
    handle = C.CDLL(_capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    for n in sorted(names):
        assert hasattr(handle, n), "libpe_b200.so does not export %s" % n
    assert names == set(_capi.PROTOTYPES), names ^ set(_capi.PROTOTYPES)
    _capi.lib()  # binds every prototype


def test_weed_layer_and_plugin_libraries_export_what_their_headers_declare():
    """include/pe_weed_layer.h (the weed_layer_t drop-ins, src/colourspace.h:387-415) and include/pe_weed_abi.h (weed_setup)"""
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(T.REPO, "include", "pe_weed_layer.h")).read(), flags=re.S)
    names = set(re.findall(r"\b([a-z][a-z0-9_]+)\s*\(", src)) - {"defined"}
    names = {n for n in names if not n.endswith("_f") and not n.endswith("_t")}  # typedefs
    assert {"convert_layer_palette_full", "convert_layer_palette", "resize_layer_full", "resize_layer", "letterbox_layer",
            "gamma_convert_layer", "gamma_convert_sub_layer", "alpha_premult", "pe_weed_layer_bind"} <= names
    handle = C.CDLL(os.path.join(T.REPO, "lives_b200", "libpe_weed_layer.so"))
    for n in sorted(names):
        assert hasattr(handle, n), "libpe_weed_layer.so does not export %s" % n
    plug = C.CDLL(os.path.join(T.REPO, "lives_b200", "libpe_weed_plugin.so"))
    assert hasattr(plug, "weed_setup") and hasattr(plug, "weed_desetup")


def test_every_entry_point_cites_the_reference():
    src = open(HEADER).read()
    for fn in ("pe_convert_layer_palette_full", "pe_resize_layer_full", "pe_letterbox_layer", "pe_gamma_convert_layer",
               "pe_gamma_convert_sub_layer", "pe_alpha_premult", "pe_fx_simple_blend", "pe_fx_multi_blend", "pe_fx_compositor"):
        i = src.index(fn + "(")
        assert re.search(r"(colourspace\.[ch]|simple_blend\.c|multi_blends\.c|compositor\.c)\s*:?\s*\d*", src[max(0, i - 700):i]), fn


def test_frame_layout_follows_reference_rowstride_rule():
    """ALIGN_CEIL(width * psize, 32); 4:2:0 / 4:2:2 chroma strides = rs0 >> 1 (colourspace.c:11299-11357)"""
    assert lb.frame_layout(1, 640, 480)[:3] == (1, [1920], [480])
    assert lb.frame_layout(3, 1280, 720)[:3] == (1, [5120], [720])
    assert lb.frame_layout(1, 37, 11)[1] == [128]
    assert lb.frame_layout(512, 1920, 1080)[:3] == (3, [1920, 960, 960], [1080, 540, 540])
    assert lb.frame_layout(522, 3840, 2160)[:3] == (3, [3840, 1920, 1920], [2160, 2160, 2160])
    assert lb.frame_layout(564, 3840, 2160)[:3] == (1, [7680], [2160])  # UYVY: 1920 macropixels x 4 bytes
    assert lb.frame_layout(545, 100, 10)[:3] == (4, [128] * 4, [10] * 4)
    with pytest.raises(ValueError):
        lb.frame_layout(9999, 10, 10)


def test_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lb.PixelEngineError) as ei:
        lb.Engine()
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under lives_b200/ may reference it"""
    root = os.path.join(T.REPO, "lives_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".c")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libpe_oracle" not in txt and "pe_or_" not in txt and "oracle/" not in txt.replace("oracle/ ", ""), f
