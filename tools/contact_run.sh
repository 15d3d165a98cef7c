for cfg in "SHELF_OVERRIDES 1048576" "PIPE_DR_OVERRIDES 1048576"; do
python tools/step_time.py $cfg +task.sim.vine_contact.binning=2
done
