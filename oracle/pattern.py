"""ORACLE helper: the rBRIEF pattern parsed from include/cmos_orb_pattern.h as a [512,2] int array."""
import os
import re

import numpy as np

_h = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "cmos_orb_pattern.h")).read()
_body = _h[_h.index("= {") + 3:_h.rindex("}")]
PATTERN = np.array([int(v) for v in re.findall(r"-?\d+", _body)], np.int32).reshape(512, 2)
