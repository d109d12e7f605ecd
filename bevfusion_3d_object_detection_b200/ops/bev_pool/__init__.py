from .bev_pool import BevPoolTables, bev_pool, bev_pool_fused

__all__ = ["bev_pool", "bev_pool_fused", "BevPoolTables"]
