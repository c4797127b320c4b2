"""Placeholder: samplers are named only inside functions of example_models/utils.py."""
