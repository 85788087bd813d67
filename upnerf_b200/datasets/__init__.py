from .ray_batcher import RayBatcher  # noqa: F401
