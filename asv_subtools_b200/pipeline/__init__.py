"""Extraction pipelines (CLI twins of pytorch/pipeline/onestep/*.py)."""
