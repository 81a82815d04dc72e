"""The one configuration value the box pipeline reads (reference data/config.py:17)."""
face = {
    'feature_maps': [160, 80, 40, 20, 10, 5],
    'min_dim': 640,
    'steps': [4, 8, 16, 32, 64, 128],
    'min_sizes': [16, 32, 64, 128, 256, 512],
    'variance': [0.1, 0.2],
    'clip': False,
    'name': 'v2',
}
