"""ORACLE - test infrastructure only.  CPU restatement (functional torch / numpy) of the
reference's distillation hot path; see dccrn_oracle.py and losses_oracle.py.  The product package
(speech-enhancement-clskd_b200/) never imports this."""
