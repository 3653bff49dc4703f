"""Mirror of the reference's ``NPHM.evaluation`` package for the part that is data-parallel: ``metrics``."""
