""" momlevel_b200 - spiciness module (mirrors ``momlevel.spice``) """

from . import flament
