#!/bin/bash
# Deferred weight gradients: which deep-level wgrads to hold back until the wide encoder levels (env sweep, one box).
bash profiles/r2_env_sweep.sh ${1:-r2zb_defer} \
  "RVIP_DEFER_WGRAD=none" \
  "RVIP_DEFER_WGRAD=dec0.conv_a,dec1.conv_a" \
  "RVIP_DEFER_WGRAD=dec0.conv_a,dec1.conv_a,dec1.upconv" \
  "RVIP_DEFER_WGRAD=dec0.conv_a,dec1.conv_a,mid.conv_b,enc3.conv_b" \
  "RVIP_DEFER_WGRAD=mid.conv_a,mid.conv_b,enc3.conv_a,enc3.conv_b,enc2.conv_a,enc2.conv_b" \
  "RVIP_DEFER_WGRAD=dec0.conv_a,dec1.conv_a RVIP_DEFER_FLUSH=enc2.conv_b" \
  "RVIP_DEFER_WGRAD=dec1.upconv,dec1.conv_a,dec1.conv_b,dec0.upconv,dec0.conv_a,dec0.conv_b,mid.conv_a,mid.conv_b,enc3.conv_a,enc3.conv_b,enc2.conv_a,enc2.conv_b" \
  "RVIP_DEFER_WGRAD=none"
