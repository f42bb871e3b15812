"""Model registry with the reference's lookup function (models/net_factory.py: get_network(name))."""
import importlib

_ROOT = __name__.split(".")[0]
_feat3dnet = importlib.import_module(("3dfeatnet_b200." if _ROOT == "3dfeatnet_b200" else "") + "models.feat3dnet")

_REGISTRY = {"3DFeatNet": _feat3dnet.Feat3dNet}


def get_network(name):
    """Class registered under `name`; KeyError for unknown names, like the reference's dict lookup."""
    return _REGISTRY[name]
