"""models/net_factory.py:3-10 of the reference: name -> model class."""
import importlib

_ROOT = __name__.split(".")[0]
Feat3dNet = importlib.import_module(("3dfeatnet_b200." if _ROOT == "3dfeatnet_b200" else "") + "models.feat3dnet").Feat3dNet

networks_map = {'3DFeatNet': Feat3dNet}


def get_network(name):
    model = networks_map[name]
    return model
