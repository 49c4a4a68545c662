"""Test infrastructure: CPU restatement of the reference hot path (see darknet_oracle.py). Never imported by the product."""
