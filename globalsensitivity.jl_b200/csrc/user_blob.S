/* Embeds the relocatable cubins of the fused kernels with an UNRESOLVED user model (user_kernel.cu.in and
 * user_kernel_separable.cu.in, built by build.py with nvcc -rdc=true -cubin) into libgsa_cuda.so;
 * gsa_model_register_fatbin / gsa_model_register_separable link them with nvJitLink.
 * Assembled once per cubin: -DGSA_BLOB_SYM=<symbol> -DGSA_BLOB_END=<symbol>_end -DGSA_USER_CUBIN_PATH="..." */
    .section .rodata
    .global GSA_BLOB_SYM
    .global GSA_BLOB_END
    .balign 16
GSA_BLOB_SYM:
    .incbin GSA_USER_CUBIN_PATH
GSA_BLOB_END:
    .section .note.GNU-stack,"",@progbits
