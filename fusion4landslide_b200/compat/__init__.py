"""Reference import paths served by the B200 hot path -- see README.md in this directory."""
