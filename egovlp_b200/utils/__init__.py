"""GPU mirrors of the reference's `utils/nDCG.py` and `utils/mAP.py` (EPIC-Kitchens MIR evaluation)."""
