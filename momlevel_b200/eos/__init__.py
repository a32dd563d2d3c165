""" momlevel_b200 - equation of state module (mirrors ``momlevel.eos``) """

from . import linear
from . import wright
