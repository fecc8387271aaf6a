"""ORACLE package: test infrastructure only (see oracle/unet_oracle.py header)."""
