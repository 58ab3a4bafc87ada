"""Stand-in for mpl_toolkits.mplot3d: ds_mppi/functions/plots.py:2 imports the name and never uses it headless."""
