"""Import-path shim: the reference's Hydra configs name classes as `src.models...`
(configs/model/*.yaml:1,16); these modules re-export the B200 implementations under those paths."""
