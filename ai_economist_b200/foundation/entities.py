"""Entity metadata (reference: entities/resources.py, landmarks.py, endogenous.py).  Only the attributes the
stepper and plotting callers read are kept: name, color, collectible / ownable / solid."""
import numpy as np

from .registrar import Registry


class Resource:
    name = None
    color = None
    collectible = None


class Landmark:
    name = None
    color = None
    ownable = None
    solid = True

    def __init__(self):
        self.blocking = self.solid and not self.ownable
        self.private = self.solid and self.ownable
        self.public = not self.solid and not self.ownable


class Endogenous:
    name = None


resource_registry = Registry(Resource)
landmark_registry = Registry(Landmark)
endogenous_registry = Registry(Endogenous)


def _resource(name, rgb, collectible):
    cls = type(name, (Resource,), dict(name=name, color=np.array(rgb) / 255.0, collectible=collectible))
    resource_registry.add(cls)
    if collectible:  # every collectible resource gets a "<R>SourceBlock" landmark (landmarks.py:55-70)
        src = type(name + "SourceBlock", (Landmark,),
                   dict(name=name + "SourceBlock", color=np.array(rgb) / 255.0, ownable=False, solid=False))
        landmark_registry.add(src)
    return cls


Wood = _resource("Wood", [107, 143, 113], True)
Stone = _resource("Stone", [241, 233, 219], True)
Coin = _resource("Coin", [229, 211, 82], False)
House = landmark_registry.add(type("House", (Landmark,), dict(name="House", color=np.array([220, 20, 220]) / 255.0,
                                                              ownable=True, solid=True)))
Water = landmark_registry.add(type("Water", (Landmark,), dict(name="Water", color=np.array([50, 50, 250]) / 255.0,
                                                              ownable=False, solid=True)))
Labor = endogenous_registry.add(type("Labor", (Endogenous,), dict(name="Labor")))
