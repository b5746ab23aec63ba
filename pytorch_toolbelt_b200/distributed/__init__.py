from .strips import StripMerger, StripPlan, exchange_halo, partition_tile_rows, sharded_merge  # noqa: F401
