"""VO parameters (ramp/config.py:3-27 defaults, overridden by config_vo/*.yaml).  The reference
uses yacs (absent here); this is a plain attribute namespace with the same field names and a
minimal `key: value` YAML reader for the five shipped config files."""
import copy


class VOConfig:
    BUFFER_SIZE = 2048
    GRADIENT_BIAS = True
    PATCHES_PER_FRAME = 80
    REMOVAL_WINDOW = 20
    OPTIMIZATION_WINDOW = 12
    PATCH_LIFETIME = 12
    KEYFRAME_INDEX = 4
    KEYFRAME_THRESH = 12.5
    MOTION_MODEL = 'DAMPED_LINEAR'
    MOTION_DAMPING = 0.5
    MIXED_PRECISION = True

    def __init__(self, **overrides):
        for k, v in overrides.items():
            if not hasattr(VOConfig, k):
                raise KeyError("unknown VO config key %r" % k)
            setattr(self, k, v)

    def clone(self):
        return copy.copy(self)

    def merge_from_file(self, path):
        with open(path) as fh:
            for line in fh:
                line = line.split("#", 1)[0].strip()
                if not line or ":" not in line:
                    continue
                k, v = [x.strip() for x in line.split(":", 1)]
                if not hasattr(VOConfig, k):
                    raise KeyError("unknown VO config key %r in %s" % (k, path))
                setattr(self, k, _parse(v))
        return self


def _parse(v):
    if v in ("True", "true"):
        return True
    if v in ("False", "false"):
        return False
    if len(v) >= 2 and v[0] == v[-1] and v[0] in "'\"":
        return v[1:-1]
    try:
        return int(v)
    except ValueError:
        try:
            return float(v)
        except ValueError:
            return v


# the shipped presets (config_vo/default.yaml, fast.yaml, precise.yaml)
PRESETS = {
    "default": dict(PATCHES_PER_FRAME=96, REMOVAL_WINDOW=22, OPTIMIZATION_WINDOW=10, PATCH_LIFETIME=13,
                    KEYFRAME_THRESH=15.0, GRADIENT_BIAS=False),
    "fast": dict(PATCHES_PER_FRAME=48, REMOVAL_WINDOW=16, OPTIMIZATION_WINDOW=7, PATCH_LIFETIME=11,
                 KEYFRAME_THRESH=15.0, GRADIENT_BIAS=False),
    "precise": dict(PATCHES_PER_FRAME=300, REMOVAL_WINDOW=42, OPTIMIZATION_WINDOW=30, PATCH_LIFETIME=33,
                    KEYFRAME_THRESH=15.0, GRADIENT_BIAS=False),
    "cfg1": dict(PATCHES_PER_FRAME=32, REMOVAL_WINDOW=22, OPTIMIZATION_WINDOW=10, PATCH_LIFETIME=13,
                 KEYFRAME_THRESH=15.0, GRADIENT_BIAS=False),
}


def preset(name):
    return VOConfig(**PRESETS[name])


cfg = VOConfig()
