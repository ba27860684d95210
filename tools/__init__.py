"""Profiling drivers and fixture generators for qdiff_b200 (not part of the product package)."""
