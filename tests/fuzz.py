"""Random node trees for parity fuzzing: every node type, random parameters inside sensible ranges, random nesting,
hybrids as constants or generators, shared sub-trees (encoded as back-references). Built as tests/encode.N descriptions so
the same tree decodes in the product, in oracle/_ref and in the C port."""
import numpy as np

from tests.encode import N, _META

LEAF_IDS = list(range(0, 12))          # no generator inputs
WARP_IDS = (16, 17, 18)
WARP_FRACTAL_IDS = (19, 20)

FLOAT_RANGES = {
    "Value": (-2.0, 2.0), "Feature Scale": (0.6, 60.0), "Output Min": (-2.0, 2.0), "Output Max": (-2.0, 2.0), "Multiplier": (-1.0, 1.0),
    "Lacunarity": (1.5, 2.6), "Amplitude Scaling": (0.4, 1.6), "Scaling": (0.3, 3.0), "Yaw": (-3.0, 3.0), "Pitch": (-3.0, 3.0), "Roll": (-3.0, 3.0),
    "Step Count": (2.0, 9.0), "Min": (-1.0, 0.0), "Max": (0.1, 1.0),
}
INT_RANGES = {"Seed Offset": (-1000, 1000), "Octaves": (1, 3), "Value Index": (0, 3), "Distance Index 0": (0, 3), "Distance Index 1": (0, 3), "Pow": (2, 5)}
HYBRID_RANGES = {
    "Gain": (0.3, 0.7), "Weighted Strength": (0.0, 0.6), "Minkowski P": (1.0, 3.0), "Grid Jitter": (0.2, 1.0), "Size Jitter": (0.0, 0.4),
    "Warp Amplitude": (0.5, 30.0), "Smoothness": (0.05, 0.5), "Ping Pong Strength": (1.5, 3.0), "Offset": (-5.0, 5.0), "Point": (-5.0, 5.0),
    "Fade": (0.0, 1.0), "Fade Min": (-1.0, 0.0), "Fade Max": (0.1, 1.0), "New Dimension Position": (-3.0, 3.0), "Pow": (0.5, 3.0),
}


class TreeFuzzer:
    def __init__(self, seed, wild=0.0):
        """wild: probability that a float parameter leaves its sensible range (zero, negative, tiny, large) and an int / enum
        member its valid one -- the trees then produce NaN, inf and signed zeros, which must follow the oracle all the same."""
        self.rng = np.random.default_rng(seed)
        self.pool = []      # finished sub-trees that may be shared
        self.wild = wild

    def uniform(self, lo, hi):
        if self.wild and self.rng.random() < self.wild:
            return float(np.float32(self.rng.choice([0.0, -0.0, -self.rng.uniform(lo, hi) if lo >= 0 else 3.0 * hi, 1e-3, 1e3, -1.0, 1.0])))
        return float(np.float32(self.rng.uniform(lo, hi)))

    def integer(self, lo, hi):
        if self.wild and self.rng.random() < self.wild:
            return int(self.rng.integers(lo - 2, hi + 3))
        return int(self.rng.integers(lo, hi + 1))

    def bounded_leaf(self):
        lo = self.uniform(0.1, 0.5)
        return N(self.rng.choice(["Perlin", "Simplex", "Value"]), {"Feature Scale": self.uniform(5.0, 40.0), "Seed Offset": int(self.rng.integers(0, 50)),
                                                                  "Output Min": lo, "Output Max": lo + self.uniform(0.1, 0.6)})

    def node(self, depth, node_id=None):
        if node_id is None:
            if depth > 0 and self.pool and self.rng.random() < 0.15:
                return self.pool[int(self.rng.integers(len(self.pool)))]  # shared sub-tree -> back-reference
            if depth <= 0:
                node_id = int(self.rng.choice(LEAF_IDS[1:] if self.rng.random() < 0.9 else LEAF_IDS))  # Constant only now and then
            else:
                node_id = int(self.rng.integers(0, 47))
                while node_id == 38:                # ConvertRGBA8 packs bytes into float BIT PATTERNS (NaNs, infinities, 1e38): below the
                    node_id = int(self.rng.integers(0, 47))  # root it would feed those to positions, where the reference itself is lane-order garbage
        meta = _META[node_id]
        variables = {}
        for v in meta["variables"]:
            if self.rng.random() < 0.3:
                continue                            # keep the default
            key = (v["name"], v["dim"]) if v["dim"] >= 0 else v["name"]
            if v["type"] == 0:
                variables[key] = self.uniform(*FLOAT_RANGES[v["name"]])
                if v["name"] == "Feature Scale" and variables[key] == 0.0:
                    variables[key] = 1.0            # a zero scale is an infinite frequency: NaN positions everywhere below
            elif v["type"] == 1:
                lo, hi = INT_RANGES[v["name"]]
                variables[key] = self.integer(lo, hi) if v["name"] != "Seed Offset" else int(self.rng.integers(lo, hi + 1))
            else:
                variables[key] = self.integer(0, len(v["enums"]) - 1)
        lookups = []
        for _ in meta["lookups"]:
            if node_id in WARP_FRACTAL_IDS:
                lookups.append(self.node(depth - 1, int(self.rng.choice(WARP_IDS))))
            else:
                lookups.append(self.node(depth - 1))
        hybrids = {}
        for position, h in enumerate(meta["hybrids"]):
            roll = self.rng.random()
            if not meta["lookups"] and position == 0 and depth > 0:
                roll = 0.5                          # a node without generator inputs gets one through its first hybrid
            if roll < 0.35:
                continue
            key = (h["name"], h["dim"]) if h["dim"] >= 0 else h["name"]
            if roll < 0.6 and depth > 0:
                hybrids[key] = self.bounded_leaf() if h["name"] in ("Gain", "Weighted Strength", "Minkowski P", "Grid Jitter", "Size Jitter", "Smoothness", "Pow") else self.node(depth - 1)
            else:
                hybrids[key] = self.uniform(*HYBRID_RANGES.get(h["name"], (-2.0, 2.0)))
        made = N(meta["name"], variables, lookups, hybrids)
        self.pool.append(made)
        return made

    def tree(self, depth=3):
        self.pool = []
        return self.node(depth, int(self.rng.integers(12, 47)))  # the root always has generator inputs


def same(a, b):
    """Bit-identical, NaNs of any payload matching NaNs."""
    a32, b32 = a.view(np.uint32), b.view(np.uint32)
    return bool(np.all((a32 == b32) | (np.isnan(a) & np.isnan(b))))
