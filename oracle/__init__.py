"""Test-infrastructure oracle package (see rvc_oracle.py header)."""
