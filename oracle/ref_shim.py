"""Load the UNMODIFIED reference sources through stub modules.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The sources are read from ``/root/reference`` (build
container) or from ``oracle/_ref`` (the vendored copy that travels to the GPU box: oracle/vendor_ref.py).  Used by
``oracle/make_golden.py`` to generate the fixtures under ``tests/golden/`` and by the
``-m "not gpu"`` tests that pin the numpy/C restatements against the real reference.

Recipe: ``tiatoolbox.tools.stainnorm`` cannot be imported normally here
because ``tiatoolbox/tools/__init__.py`` pulls in umap/skimage/... which are absent.  We
register bare namespace packages for ``tiatoolbox.utils`` / ``tiatoolbox.tools`` /
``tiatoolbox.models[.architecture]`` (only ``__path__``), synthesise
``tiatoolbox.utils.misc`` from the *verbatim text* of the handful of functions the hot
path uses (AST-extracted from ``utils/misc.py``), and then let importlib load
``stainnorm.py``, ``stainextract.py``, ``utils/transforms.py``, ``hovernet.py``,
``unet.py``, ``vanilla.py`` unmodified.

The only restated third-party function is ``skimage.exposure.rescale_intensity``
(scikit-image is not installed); the restatement is pinned by the reference's own
golden vector ``tests/test_utils.py:882-911``.
"""
from __future__ import annotations

import ast
import importlib
import sys
import types
from pathlib import Path

import numpy as np

# /root/reference in the build container; on the GPU box the byte-for-byte copy of the hot-path files that oracle/vendor_ref.py
# places under oracle/_ref/ (git-ignored, travels with the snapshot like the built .so)
_CANDIDATES = (Path("/root/reference"), Path(__file__).resolve().parent / "_ref")
REF_ROOT = next((c for c in _CANDIDATES if (c / "tiatoolbox" / "tools" / "stainnorm.py").exists()), _CANDIDATES[0])
REF_PKG = REF_ROOT / "tiatoolbox"


def available() -> bool:
    return (REF_PKG / "tools" / "stainnorm.py").exists()


class _Exposure:
    """Restatement of skimage.exposure.rescale_intensity for (uint8 image, float in_range,
    float out_range) — the only way the reference calls it (utils/misc.py:439-443)."""

    @staticmethod
    def rescale_intensity(image, in_range, out_range):
        imin, imax = float(in_range[0]), float(in_range[1])
        omin, omax = float(out_range[0]), float(out_range[1])
        image = np.clip(image.astype(np.float64), imin, imax)
        if imin != imax:
            image = (image - imin) / (imax - imin)
            return image * (omax - omin) + omin
        return np.clip(image, omin, omax)


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_MISC_WANTED = [
    "contrast_enhancer",
    "get_luminosity_tissue_mask",
    "load_stain_matrix",
    "split_path_name_ext",
    "get_bounding_box",
    "parse_cv2_interpolaton",
    "select_cv2_interpolation",
]

_loaded = {}


def _base_packages():
    """Register tiatoolbox + bare namespace sub-packages (skipping their __init__)."""
    if "base" in _loaded:
        return
    if str(REF_ROOT) not in sys.path:
        # appended, not prepended: /root/reference has a top-level `tests` package of its own that must not shadow this repository's
        # (spawned gloo workers re-import `tests.<module>` with the parent's sys.path)
        sys.path.append(str(REF_ROOT))
    import tiatoolbox  # noqa: F401  (top-level import is light: yaml + logging)

    for sub in ("utils", "tools", "models"):
        m = _stub(f"tiatoolbox.{sub}")
        m.__path__ = [str(REF_PKG / sub)]
    m = _stub("tiatoolbox.models.architecture")
    m.__path__ = [str(REF_PKG / "models" / "architecture")]
    importlib.import_module("tiatoolbox.utils.exceptions")  # dependency-free
    _loaded["base"] = True


def _extract_misc():
    import cv2
    import pandas as pd

    _base_packages()
    from tiatoolbox.utils.exceptions import FileNotSupportedError

    src = (REF_PKG / "utils" / "misc.py").read_text()
    fns = [
        ast.get_source_segment(src, n)
        for n in ast.parse(src).body
        if isinstance(n, ast.FunctionDef) and n.name in _MISC_WANTED
    ]
    ns = {
        "np": np,
        "cv2": cv2,
        "pd": pd,
        "Path": Path,
        "exposure": _Exposure,
        "FileNotSupportedError": FileNotSupportedError,
    }
    exec("from __future__ import annotations\n" + "\n\n".join(fns), ns)  # noqa: S102
    return {k: ns[k] for k in _MISC_WANTED}


def load_stain():
    """Return the reference's ``tiatoolbox.tools.stainnorm`` module (unmodified source)."""
    if "stain" in _loaded:
        return _loaded["stain"]
    if not available():
        raise RuntimeError("reference not present at /root/reference")
    misc = _extract_misc()
    _stub("tiatoolbox.utils.misc", **misc)
    sn = importlib.import_module("tiatoolbox.tools.stainnorm")
    _loaded["stain"] = sn
    return sn


def load_misc():
    load_stain()
    return sys.modules["tiatoolbox.utils.misc"]


def target_image() -> np.ndarray:
    """The in-tree 1000x1000x3 stain target (tiatoolbox/data/target_image.png)."""
    import cv2

    bgr = cv2.imread(str(REF_PKG / "data" / "target_image.png"), cv2.IMREAD_COLOR)
    return np.ascontiguousarray(bgr[..., ::-1])


def load_models():
    """Return the reference's architecture modules (hovernet, unet, vanilla, utils) loaded
    unmodified, with the absent third-party packages (dask, skimage, timm) stubbed out."""
    if "models" in _loaded:
        return _loaded["models"]
    load_stain()  # base packages + tiatoolbox.utils.misc stub (get_bounding_box included)

    class _Array:  # dask.array.Array placeholder for type annotations
        pass

    if "dask" not in sys.modules:
        da = _stub("dask.array", Array=_Array)
        d = _stub("dask", array=da, delayed=lambda f: f, compute=lambda *a, **k: a)
        d.__path__ = []
    for name in ("skimage", "skimage.morphology", "skimage.segmentation", "skimage.feature", "skimage.measure", "skimage.exposure"):
        if name not in sys.modules:
            m = _stub(name)
            m.__path__ = []
    sys.modules["skimage.morphology"].remove_small_objects = None
    sys.modules["skimage.segmentation"].watershed = None
    sys.modules["skimage.feature"].peak_local_max = None
    sys.modules["skimage.measure"].label = None
    sys.modules["skimage.measure"].regionprops = None
    if "timm" not in sys.modules:
        t = _stub("timm", list_models=lambda *a, **k: [], create_model=None)
        t.__path__ = []
        t.layers = _stub("timm.layers", SwiGLUPacked=None)
        t.models = _stub("timm.models", VisionTransformer=None)
    mods = types.SimpleNamespace()
    mods.models_abc = importlib.import_module("tiatoolbox.models.models_abc")
    mods.utils = importlib.import_module("tiatoolbox.models.architecture.utils")
    mods.hovernet = importlib.import_module("tiatoolbox.models.architecture.hovernet")
    mods.unet = importlib.import_module("tiatoolbox.models.architecture.unet")
    mods.vanilla = importlib.import_module("tiatoolbox.models.architecture.vanilla")
    _loaded["models"] = mods
    return mods


def load_semseg_merge():
    """The reference's pure-numpy stitching helpers of tiatoolbox/models/engine/semantic_segmentor.py, executed from their
    unmodified source text (the module itself imports dask / zarr / psutil, which this image does not have):
    merge_batch_to_canvas :1141-1183, clip_probabilities_to_shape :1537-1557, concatenate_none :1112-1138 (numpy branch)."""
    if "semseg_merge" in _loaded:
        return _loaded["semseg_merge"]
    src = (REF_PKG / "models" / "engine" / "semantic_segmentor.py").read_text()
    wanted = {"merge_batch_to_canvas", "clip_probabilities_to_shape"}
    code = "from __future__ import annotations\nimport numpy as np\n" + "\n\n".join(
        ast.get_source_segment(src, n)
        for n in ast.parse(src).body
        if isinstance(n, ast.FunctionDef) and n.name in wanted
    )
    mod = types.ModuleType("tiatoolbox_ref_semseg_merge")
    exec(compile(code, "semantic_segmentor.py[merge helpers]", "exec"), mod.__dict__)  # noqa: S102
    _loaded["semseg_merge"] = mod
    return mod


class _Box:
    """Restatement of the only shapely geometry the tile merge uses: an axis-aligned `shapely.box(minx, miny, maxx, maxy)` (shapely is a
    pinned but absent dependency of the reference; `requirements/requirements.txt`).  `contains` follows the DE-9IM definition for
    rectangles: every point of `other` lies in self and the interiors meet, i.e. closed-interval inclusion in a self that has area."""

    def __init__(self, minx, miny, maxx, maxy):
        self.b = (min(minx, maxx), min(miny, maxy), max(minx, maxx), max(miny, maxy))

    def contains(self, other: "_Box") -> bool:
        a, o = self.b, other.b
        if a[2] <= a[0] or a[3] <= a[1]:
            return False   # a line / point has no interior that a rectangle could share
        inside = a[0] <= o[0] and a[1] <= o[1] and o[2] <= a[2] and o[3] <= a[3]
        return bool(inside)


class _STRtree:
    """`STRtree(geoms).query(geom)` without a predicate: the indices of the geometries whose envelopes intersect the envelope of `geom`
    (closed intervals: touching counts).  shapely returns them in tree order; every call site in the reference only uses them as a set."""

    def __init__(self, geoms):
        self.bounds = np.array([g.b for g in geoms], dtype=np.float64).reshape(-1, 4)

    def query(self, geom):
        q = geom.b
        b = self.bounds
        hit = (b[:, 0] <= q[2]) & (b[:, 2] >= q[0]) & (b[:, 1] <= q[3]) & (b[:, 3] >= q[1])
        return np.nonzero(hit)[0]


_WSI_MERGE_FUNCS = [
    "_process_instance_predictions", "retrieve_sel_uids", "_get_sel_indices_margin_lines", "_get_margin_lines",
    "_move_tile_space_to_wsi_space", "_get_inst_info_dicts", "_update_tile_based_predictions_array", "_build_tile_tasks",
    "_compute_info_dict_for_merge",
]


def load_wsi_merge():
    """The reference's WSI tile-mode merging of tiatoolbox/models/engine/multi_task_segmentor.py, executed from the unmodified source
    text of its functions (the module imports dask / zarr / shapely / matplotlib / psutil, none of which this image has):
    `_get_tile_info` :1362-1554 (a method: loaded as a plain function taking `self`), `_process_instance_predictions` :2833-2933 and
    its helpers :2936-3058, `_get_inst_info_dicts` :3061, `_update_tile_based_predictions_array` :3160-3212, `_build_tile_tasks` :3215,
    `_compute_info_dict_for_merge` :3269; plus `PatchExtractor.get_coordinates` (tools/patchextraction.py:488-613) and `validate_shape`.
    shapely's `box` / `STRtree` are the restatements above."""
    if "wsi_merge" in _loaded:
        return _loaded["wsi_merge"]
    import textwrap
    import uuid
    from collections import deque

    mod = types.ModuleType("tiatoolbox_ref_wsi_merge")
    pe_src = (REF_PKG / "tools" / "patchextraction.py").read_text()
    pe_tree = ast.parse(pe_src)
    parts = ["from __future__ import annotations", "import numpy as np"]
    for n in pe_tree.body:
        if isinstance(n, ast.FunctionDef) and n.name == "validate_shape":
            parts.append(ast.get_source_segment(pe_src, n))
        if isinstance(n, ast.ClassDef) and n.name == "PatchExtractor":
            for m in n.body:
                if isinstance(m, ast.FunctionDef) and m.name == "get_coordinates":
                    parts.append("class PatchExtractor:\n    @staticmethod\n" + textwrap.indent(textwrap.dedent(
                        "    " * 0 + ast.get_source_segment(pe_src, m, padded=True)), "    "))
    src = (REF_PKG / "models" / "engine" / "multi_task_segmentor.py").read_text()
    tree = ast.parse(src)
    for n in tree.body:
        if isinstance(n, ast.FunctionDef) and n.name in _WSI_MERGE_FUNCS:
            parts.append(ast.get_source_segment(src, n))
        if isinstance(n, ast.ClassDef) and n.name == "MultiTaskSegmentor":
            for m in n.body:
                if isinstance(m, ast.FunctionDef) and m.name == "_get_tile_info":
                    parts.append(textwrap.dedent(ast.get_source_segment(src, m, padded=True)))
    shapely = types.SimpleNamespace(box=_Box)
    mod.__dict__.update({"shapely": shapely, "STRtree": _STRtree, "uuid": uuid, "deque": deque,
                         "tqdm": lambda it, **_k: it})
    exec(compile("\n\n".join(parts), "multi_task_segmentor.py[wsi merge]", "exec"), mod.__dict__)  # noqa: S102
    _loaded["wsi_merge"] = mod
    return mod


def load_stainaugment():
    """The reference's `tiatoolbox.tools.stainaugment` module, unmodified, with albumentations' `ImageOnlyTransform` base class (absent
    here) replaced by a two-argument stand-in — the reference only uses its constructor (stainaugment.py:111)."""
    if "stainaugment" in _loaded:
        return _loaded["stainaugment"]
    load_stain()

    class ImageOnlyTransform:
        def __init__(self, always_apply=False, p=0.5):
            self.always_apply, self.p = always_apply, p

    if "albumentations" not in sys.modules:
        a = _stub("albumentations")
        a.__path__ = []
        c = _stub("albumentations.core")
        c.__path__ = []
        _stub("albumentations.core.transforms_interface", ImageOnlyTransform=ImageOnlyTransform)
    mod = importlib.import_module("tiatoolbox.tools.stainaugment")
    _loaded["stainaugment"] = mod
    return mod
