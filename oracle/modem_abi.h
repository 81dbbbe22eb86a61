/* oracle/modem_abi.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Force-included (-include) in front of the reference's OWN, unmodified sources -- the modems
 * (/root/reference/src/modules/modem/Modem.cpp, ModemAnalog.cpp, analog/Modem*.cpp) and the visual processors
 * (src/process/FFTDataDistributor.cpp, SpectrumVisualProcessor.cpp, ScopeVisualProcessor.cpp) -- when oracle/build_ref.py
 * compiles them where they lie into oracle/_ref/libmodems_ref.so / libvisuals_ref.so.  Those sources call liquid-dsp through liquid.h's System V
 * prototypes; the only liquid binary the reference ships is a MinGW archive whose functions use the Microsoft x64
 * convention (oracle/build_ref.py).  Every liquid entry point the modem sources use is therefore renamed here to a
 * System V forwarder sv_<name> (oracle/ref_modem_abi.c).  liquid.h builds its names by token pasting and the pasted
 * token is rescanned, so the renaming reaches the declarations in liquid.h and the calls in the modem sources alike.
 */
#ifndef ORACLE_MODEM_ABI_H
#define ORACLE_MODEM_ABI_H
#define ampmodem_create sv_ampmodem_create
#define ampmodem_demodulate sv_ampmodem_demodulate
#define ampmodem_destroy sv_ampmodem_destroy
#define estimate_req_filter_len sv_estimate_req_filter_len
#define firfilt_rrrf_create sv_firfilt_rrrf_create
#define firfilt_rrrf_create_dc_blocker sv_firfilt_rrrf_create_dc_blocker
#define firfilt_rrrf_destroy sv_firfilt_rrrf_destroy
#define firfilt_rrrf_execute sv_firfilt_rrrf_execute
#define firfilt_rrrf_push sv_firfilt_rrrf_push
#define firhilbf_c2r_execute sv_firhilbf_c2r_execute
#define firhilbf_create sv_firhilbf_create
#define firhilbf_destroy sv_firhilbf_destroy
#define firhilbf_r2c_execute sv_firhilbf_r2c_execute
#define freqdem_create sv_freqdem_create
#define freqdem_demodulate_block sv_freqdem_demodulate_block
#define freqdem_destroy sv_freqdem_destroy
#define iirfilt_crcf_create_lowpass sv_iirfilt_crcf_create_lowpass
#define iirfilt_crcf_create_prototype sv_iirfilt_crcf_create_prototype
#define iirfilt_crcf_destroy sv_iirfilt_crcf_destroy
#define iirfilt_crcf_execute sv_iirfilt_crcf_execute
#define iirfilt_rrrf_create sv_iirfilt_rrrf_create
#define iirfilt_rrrf_destroy sv_iirfilt_rrrf_destroy
#define iirfilt_rrrf_execute sv_iirfilt_rrrf_execute
#define liquid_firdes_kaiser sv_liquid_firdes_kaiser
#define msresamp_cccf_create sv_msresamp_cccf_create
#define msresamp_cccf_destroy sv_msresamp_cccf_destroy
#define msresamp_cccf_execute sv_msresamp_cccf_execute
#define msresamp_rrrf_create sv_msresamp_rrrf_create
#define msresamp_rrrf_destroy sv_msresamp_rrrf_destroy
#define msresamp_rrrf_execute sv_msresamp_rrrf_execute
#define nco_crcf_cexpf sv_nco_crcf_cexpf
#define nco_crcf_create sv_nco_crcf_create
#define nco_crcf_destroy sv_nco_crcf_destroy
#define nco_crcf_get_frequency sv_nco_crcf_get_frequency
#define nco_crcf_mix_down sv_nco_crcf_mix_down
#define nco_crcf_mix_up sv_nco_crcf_mix_up
#define nco_crcf_pll_set_bandwidth sv_nco_crcf_pll_set_bandwidth
#define nco_crcf_pll_step sv_nco_crcf_pll_step
#define nco_crcf_reset sv_nco_crcf_reset
#define nco_crcf_set_frequency sv_nco_crcf_set_frequency
#define nco_crcf_step sv_nco_crcf_step
/* the visual processors (FFTDataDistributor needs none; SpectrumVisualProcessor.cpp, ScopeVisualProcessor.cpp) */
#define fft_create_plan sv_fft_create_plan
#define fft_execute sv_fft_execute
#define fft_destroy_plan sv_fft_destroy_plan
#define nco_crcf_mix_block_up sv_nco_crcf_mix_block_up
#define nco_crcf_mix_block_down sv_nco_crcf_mix_block_down
#define msresamp_crcf_create sv_msresamp_crcf_create
#define msresamp_crcf_destroy sv_msresamp_crcf_destroy
#define msresamp_crcf_execute sv_msresamp_crcf_execute
/* the ingress (src/sdr/SDRPostThread.cpp) */
#define firpfbch2_crcf_create_kaiser sv_firpfbch2_crcf_create_kaiser
#define firpfbch2_crcf_destroy sv_firpfbch2_crcf_destroy
#define firpfbch2_crcf_execute sv_firpfbch2_crcf_execute
#define firpfbch_crcf_analyzer_execute sv_firpfbch_crcf_analyzer_execute
#define firpfbch_crcf_create_kaiser sv_firpfbch_crcf_create_kaiser
#define firpfbch_crcf_destroy sv_firpfbch_crcf_destroy
#define iirfilt_crcf_create_dc_blocker sv_iirfilt_crcf_create_dc_blocker
#define iirfilt_crcf_execute_block sv_iirfilt_crcf_execute_block
#endif
