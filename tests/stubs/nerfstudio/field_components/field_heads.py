from enum import Enum


class FieldHeadNames(Enum):
    RGB = "rgb"
    DEPTH = "depth"
    NORMALS = "normals"
