# Dev tool: far-field opening parameters vs error and time (scripts/time_contact.py, B=256)
for bl in 1.4 1.6 1.8 2.0; do for bs in 2.0 2.5; do
  echo "beta_leaf=$bl beta_group=$bs"; TUCH_WC_BETA=$bl TUCH_WC_BETA_SUPER=$bs python scripts/time_contact.py 256 2>&1 | grep -E "fast winding|nearest: tiles"
done; done
