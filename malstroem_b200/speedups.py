"""The plug-in switch, with the contract of malstroem/algorithms/speedups/__init__.py:22-100:
`available`, `enabled`, `enable()`, `disable()`.

enable() rebinds the twelve whole-function names the tool layer resolves at call time
(malstroem/dem.py:67-89, malstroem/bluespots.py:159-205, malstroem/scripts/dem.py, scripts/bluespot.py) on the
reference's own `malstroem.algorithms.{fill,flow,label}` modules to the B200 implementations, and leaves
`malstroem.algorithms.speedups.enabled` truthy so the tools do not warn (dem.py:62, bluespots.py:154).
Beyond the twelve: the §8(f) rows that are built (pour-point network, Network.rain_event, vectorize_labels_file).
Nothing else of the reference is touched; disable() restores the saved originals.
"""
import os
import warnings

from . import _lib
import importlib

from . import network as _network
from . import vector as _vector
from .algorithms import fill as _fill, flow as _flow, label as _label, net as _net

available = os.path.exists(_lib.LIB_PATH)
enabled = False
_orig = {}

REBOUND = {
    "fill": ("fill_terrain", "fill_terrain_no_flats", "minimum_safe_short_and_diag"),
    "flow": ("terrain_flowdirection", "accumulated_flow", "watersheds_from_labels"),
    "label": ("connected_components", "label_stats", "keep_labels", "label_min_index", "label_max_index",
              "label_count"),
}
# the pour-point network functions StreamTool resolves at call time (streams.py:72-74)
REBOUND_NEXT = {
    "net": ("next_downstream_label", "pourpoint_network", "geometric_pourpoint_network"),
}
_OURS = {"fill": _fill, "flow": _flow, "label": _label, "net": _net}


def enable(target=None):
    """Rebind the hot-path functions of `target` (default: the imported `malstroem.algorithms` package)."""
    global enabled
    if not available:
        warnings.warn("malstroem_b200: libmalstroem_b200.so not built; nothing enabled", RuntimeWarning)
        return
    if _orig:
        return
    if target is None:
        import malstroem.algorithms as target      # the reference package must be importable
    for modname, names in REBOUND.items():
        mod = getattr(target, modname)
        for name in names:
            _orig[(modname, name)] = (mod, getattr(mod, name))
            setattr(mod, name, getattr(_OURS[modname], name))
    for modname, names in REBOUND_NEXT.items():
        mod = getattr(target, modname, None)
        if mod is None:
            try:
                mod = importlib.import_module(target.__name__ + "." + modname)
            except ImportError:
                continue
        for name in names:
            _orig[(modname, name)] = (mod, getattr(mod, name))
            setattr(mod, name, getattr(_OURS[modname], name))
    # RainTool holds the class itself (rain.py:17), so the method is replaced on the class;
    # Network.rain_event below only needs the instance's `nodes` list, which the reference class keeps too
    try:
        ref_network = importlib.import_module(target.__name__.split(".")[0] + ".network")
        cls = ref_network.Network
        _orig[("network", "Network.rain_event")] = (cls, cls.__dict__["rain_event"])
        cls.rain_event = _rain_event_on_reference_instance
    except (ImportError, AttributeError, KeyError):
        pass
    # BluespotTool binds vectorize_labels_file by name at import (bluespots.py:17, used at
    # bluespots.py:179,192), so it is replaced on both modules
    for refmod in ("vector", "bluespots"):
        try:
            mod = importlib.import_module(target.__name__.split(".")[0] + "." + refmod)
            _orig[(refmod, "vectorize_labels_file")] = (mod, mod.__dict__["vectorize_labels_file"])
            mod.vectorize_labels_file = _vector.vectorize_labels_file
        except (ImportError, KeyError):
            pass
    ref_speedups = getattr(target, "speedups", None)
    if ref_speedups is not None:
        _orig[("speedups", "enabled")] = (ref_speedups, ref_speedups.enabled)
        ref_speedups.enabled = True
    enabled = True


def _rain_event_on_reference_instance(self, mmrain):
    """Network.rain_event (network.py:113-129) for an instance of the REFERENCE's class."""
    ours = _network.Network.__new__(_network.Network)
    ours.nodes = self.nodes
    events = _network.Network.rain_events(ours, [mmrain])[0]
    self._node_rain_values = {e['nodeid']: e for e in events}
    return events


def disable():
    global enabled
    if not _orig:
        return
    for (modname, name), (mod, fn) in _orig.items():
        setattr(mod, name.split(".")[-1], fn)
    _orig.clear()
    enabled = False
