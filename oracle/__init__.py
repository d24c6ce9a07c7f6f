"""oracle/: CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY --
see the header of knockoffs_oracle.py.  The product package never imports this."""
