; bsgs_b200.pbi -- PureBasic binding of include/bsgs_b200.h (the C ABI of libbsgs_b200.so).
;
; IncludeFile this from 1_9_7File.pb in place of the Import "lib\cuda.lib" block (1_9_7File.pb:55-106)
; and replace the body of cuda() as INTEGRATION.md section 2 shows.  NOT COMPILED HERE: this image has
; no PureBasic compiler; the same entry points, with the same argument order, are exercised by the
; C++ host (bsgs-cuda_b200/csrc/host_main.cpp) and by the ctypes binding (bsgs-cuda_b200/native.py).
; All big integers are the 32-byte little-endian buffers of lib\Curve64.pb (m_sethex32, 450-473).

#BSGS_OK = 0
#BSGS_E_ARG = -1
#BSGS_E_CUDA = -2
#BSGS_E_NOMEM = -3
#BSGS_E_STATE = -4
#BSGS_E_LIMIT = -5
#BSGS_E_OVERFLOW = -6
#BSGS_MAX_WINDOWS = 16

Structure bsgs_hit            ; one record: code 1/2/4/5 and idx as written by the reference kernel
  code.l                      ;   (PeekL(*a+8*i), PeekL(*a+8*i+4), 1_9_7File.pb:2470-2490)
  idx.l
  window.l                    ; which of the windows of this call (0 for bsgs_step)
  reserved.l                  ; with option "verify": payload j of the table entry (0xFFFFFFFE for code 4)
EndStructure

Structure bsgs_devinfo
  name.a[128]
  cc_major.l : cc_minor.l : sm_count.l : reserved.l
  free_bytes.q : total_bytes.q
EndStructure

CompilerIf #PB_Compiler_OS = #PB_OS_Windows
  Import "bsgs_b200.lib"
CompilerElse
  ImportC "libbsgs_b200.so"
CompilerEndIf
  bsgs_create(*ctx, device.l, T.l, B.l, p.l, w.q, htsz.l)              ; cuInit .. cuMemAlloc, 2181-2255
  bsgs_destroy(*ctx)                                                   ; cuMemFree / cuCtxDestroy, 2537-2548
  bsgs_strerror(status.l)
  bsgs_last_cuda_error(*ctx)
  bsgs_set_option(*ctx, name.p-ascii, value.q)                         ; "verify", "overlap", "blocks_per_sm", "block", "max_windows", "l2_fetch"
  bsgs_upload_giants(*ctx, *g2, bytes.i)                               ; cuMemcpyHtoD of *GiantArrPacked, 2337
  bsgs_upload_ht(*ctx, *htgpu, bytes.i)                                ; cuMemcpyHtoD of *GpuHT, 2350
  bsgs_upload_htcpu(*ctx, *htcpu, bytes.i)                             ; same table with payloads: device-side verification
  bsgs_bcast_ht(*ctxs, n.l)                                            ; table of ctxs(0) -> all others, device to device
  bsgs_giants_generate(*ctx)                                           ; instead of Save_Load_Giants, 1905-2058
  bsgs_giants_export(*ctx, *g2, bytes.i)
  bsgs_ht_generate(*ctx)                                               ; instead of GenBabys + sort + pack, 1237-1328, 3645-3759
  bsgs_ht_export(*ctx, kind.l, *out, bytes.i)                          ; kind 0: _htGPUv0.BIN image, 1: _htCPUv0.BIN image
  bsgs_step(*ctx, *px, *py, *hits, cap.l, *nhits)                      ; one launch, 2435-2509
  bsgs_step_multi(*ctx, nwin.l, *pxy, *hits, cap.l, *nhits)
  bsgs_run(*ctx, *p0xy, *dxy, nwindows.q, batch.l, stop_after_hits.l, *hits, cap.l, *nhits, *windows_done)   ; GetJob loop; two launches in flight when nwindows > batch
  bsgs_device_count(*count)                                            ; start-up listing + Tune, 778-860, 324-431
  bsgs_device_query(device.l, *devinfo)
EndImport
