"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference path (see gsdr_oracle.py)."""
