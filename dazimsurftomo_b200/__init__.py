"""B200-native forward-modelling path of DAzimSurfTomo (depth kernels -> eikonal -> rays -> G rows)."""
