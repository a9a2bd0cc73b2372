"""Constants re-exported to user scripts (reference sdf/util.py:1-6)."""
import math

pi = math.pi

degrees = math.degrees
radians = math.radians
