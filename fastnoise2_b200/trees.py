"""Encoded node trees of the BASELINE.json configs. Pure data: importing this module loads no native code (bench.py's
reference arm and the oracle-side tests read it without mapping libfastnoise2_b200.so)."""

# Encoded trees of the five BASELINE.json configs (each decodes with the reference's own decoder)
TREES = {
    "fbm_simplex_readme": "DQkGDA==",          # FastNoise.h:53 example: FractalFBm( Simplex )
    "c1_simplex": "BgQ=",                      # config 1: Simplex, 3D grid 256^3
    "c2_fbm5_perlin": "DQAFAAAACQgM",          # config 2: FractalFBm( 5 octaves, Perlin ), 2D grid 16384^2 / tileable 4096^2
    "c3a_cellular_distance": "CwQ=",           # config 3: CellularDistance, 3D grid 512^3
    "c3b_cellular_lookup_white": "DAkBDA==",   # config 3: CellularLookup( White ): output equal <=> same cell, bit for bit
    "c3c_cellular_lookup_gradient": "DAkEAArXIzwI93VgPBCad5w8DA==",
    "c4_terrain": "EwkSAAAASEMJKgAAAEBBCSkAAQAAAAkdCQ0ABQAAAAkGAAAAFkMMAg8ABAAAAAkIAAAAtEIIBwAAAAwLzcxMPgQLAAAAQAQDAACAPgQDAADwQQw=",
    "c5_supersimplex": "BwQ=",                 # config 5: SuperSimplex, 4D position array
}
