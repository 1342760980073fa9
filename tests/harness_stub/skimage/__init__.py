"""Test-harness stand-in for scikit-image (not installed here): base_trainer.py:16 imports two metrics."""
