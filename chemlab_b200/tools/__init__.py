"""Command-line table tools that ship next to chemlab's driver (reference: tools/): run as `python -m chemlab_b200.tools.<name>`."""
