"""Link lengths of the planar arms (reference: utils/constants.py:1-13 -- every link is 1.0 long)."""


class Constants:
  # two-link arm
  R2_LENGTH1 = 1.0
  R2_LENGTH2 = 1.0
  # three-link arm
  R3_LENGTH1 = 1.0
  R3_LENGTH2 = 1.0
  R3_LENGTH3 = 1.0
  # eight-link arm (never runnable in the reference)
  R8_LENGTH = 1.0
