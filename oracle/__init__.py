"""ORACLE -- test infrastructure only. See oracle/codec_oracle.py and DESIGN.md."""
