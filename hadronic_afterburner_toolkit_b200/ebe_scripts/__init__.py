"""Downstream consumers of the (frozen) HBT output files — SURVEY.md §8f rank 4.

Host-side numpy/scipy only: they read what `output_correlation_function` wrote
(`/root/reference/src/HBT_correlation.cpp:726-783`), whichever implementation wrote it."""
