/* Embeds the relocatable cubin of the generic fused kernel with an unresolved model (user_kernel.cu.in, built by
 * build.py with nvcc -rdc=true -cubin) into libgsa_cuda.so; gsa_model_register_fatbin links it with nvJitLink. */
    .section .rodata
    .global gsa_user_kernel_blob
    .global gsa_user_kernel_blob_end
    .balign 16
gsa_user_kernel_blob:
    .incbin GSA_USER_CUBIN_PATH
gsa_user_kernel_blob_end:
    .section .note.GNU-stack,"",@progbits
