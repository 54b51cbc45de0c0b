"""Test infrastructure only: CPU oracle of the wav2sleep forward (see wav2sleep_oracle.py)."""
