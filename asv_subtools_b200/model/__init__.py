"""Model blueprints (one class per file, loaded by path through `create_model_from_py`, like
pytorch/model/*.py in the reference)."""
