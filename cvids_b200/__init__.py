"""B200-native OpenChisel hot path (TSDF integration + incremental marching cubes) for CVIDS."""
