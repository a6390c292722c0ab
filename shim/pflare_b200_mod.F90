! pflare_b200_mod.F90 -- Fortran side of the pflare-b200 binding (add to PFLARE's src/).
!
! NOT COMPILED IN THIS REPOSITORY (no gfortran / PETSc here).  Handle conventions follow
! src/C_PETSc_Interfaces.F90:156-198 (opaque type(c_ptr) contexts passed by reference for create/destroy) and
! src/C_Fortran_Bindings.F90:48-64 (PETSc objects cross the boundary as integer(c_long_long) = obj%v).
module pflare_b200_mod

   use iso_c_binding
   use petscmat
   use air_data_type
   use matshell_data_type
   use pflare_parameters

#include "petsc/finclude/petscmat.h"

   implicit none
   public

   ! operator selectors of include/pflare_b200.h
   integer(c_int), parameter :: B200_AFF = 0, B200_AFC = 1, B200_ACF = 2, B200_ACC = 3, B200_INV_AFF = 4, &
                                B200_INV_ACC = 5, B200_R = 6, B200_P = 7, B200_COARSE = 8

   interface
      subroutine pflare_b200_create_c(handle, aux, A_top, no_levels) bind(c, name="pflare_b200_create_c")
         use iso_c_binding
         type(c_ptr) :: handle, aux                    ! aux: per-PC communicator slot + ordering events (air_data%b200_aux)
         integer(c_long_long) :: A_top
         integer(c_long_long), value :: no_levels      ! PetscInt
      end subroutine
      subroutine pflare_b200_set_level_c(handle, our_level, A_level, is_fine, is_coarse, smooth_order, n_smooth) &
                  bind(c, name="pflare_b200_set_level_c")
         use iso_c_binding
         type(c_ptr) :: handle
         integer(c_long_long), value :: our_level
         integer(c_long_long) :: A_level, is_fine, is_coarse
         integer(c_int), dimension(*) :: smooth_order
         integer(c_int), value :: n_smooth
      end subroutine
      subroutine pflare_b200_upload_mat_c(handle, our_level, which, A) bind(c, name="pflare_b200_upload_mat_c")
         use iso_c_binding
         type(c_ptr) :: handle
         integer(c_long_long), value :: our_level
         integer(c_int), value :: which
         integer(c_long_long) :: A
      end subroutine
      subroutine pflare_b200_upload_diag_c(handle, our_level, which, D) bind(c, name="pflare_b200_upload_diag_c")
         use iso_c_binding
         type(c_ptr) :: handle
         integer(c_long_long), value :: our_level
         integer(c_int), value :: which
         integer(c_long_long) :: D
      end subroutine
      function pflare_b200_set_poly(handle, our_level, which, inverse_type, ncoef, coeffs_re, coeffs_im, diag_scale) &
                  bind(c, name="pflare_b200_set_poly")
         use iso_c_binding
         type(c_ptr), value :: handle
         integer(c_int), value :: our_level, which, inverse_type, ncoef, diag_scale
         real(c_double), dimension(*) :: coeffs_re, coeffs_im
         integer(c_int) :: pflare_b200_set_poly
      end function
      subroutine pflare_b200_finalize_c(handle) bind(c, name="pflare_b200_finalize_c")
         use iso_c_binding
         type(c_ptr) :: handle
      end subroutine
      subroutine pflare_b200_apply_c(handle, aux, x, y) bind(c, name="pflare_b200_apply_c")
         use iso_c_binding
         type(c_ptr) :: handle, aux
         integer(c_long_long) :: x, y
      end subroutine
      subroutine pflare_b200_destroy_c(handle, aux) bind(c, name="pflare_b200_destroy_c")
         use iso_c_binding
         type(c_ptr) :: handle, aux
      end subroutine
      subroutine pflare_b200_coarse_its_c(pcmg, its) bind(c, name="pflare_b200_coarse_its_c")
         use iso_c_binding
         integer(c_long_long) :: pcmg
         integer(c_int) :: its
      end subroutine
      function pflare_b200_set_option(handle, key, val) bind(c, name="pflare_b200_set_option")
         use iso_c_binding
         type(c_ptr), value :: handle
         character(kind=c_char), dimension(*) :: key
         real(c_double), value :: val
         integer(c_int) :: pflare_b200_set_option
      end function
   end interface

   contains

   ! One approximate inverse -> set_csr | set_diag | set_poly, following the dispatch of
   ! src/Approx_Inverse_Setup.F90:394-500 and the MatShell contexts of src/Gmres_Poly_Newton.F90:1993-2010
   subroutine upload_inverse_b200(handle, our_level, which, inv_mat, inverse_type, diag_scale)
      type(c_ptr), intent(inout) :: handle
      integer, intent(in)        :: our_level, which, inverse_type
      type(tMat), intent(in)     :: inv_mat
      logical, intent(in)        :: diag_scale

      MatType :: mat_type
      type(mat_ctxtype), pointer :: mat_ctx
      PetscErrorCode :: ierr
      integer(c_int) :: ierr_c
      real(c_double), dimension(1) :: dummy_im

      call MatGetType(inv_mat, mat_type, ierr)
      if (mat_type == MATDIAGONAL) then
         call pflare_b200_upload_diag_c(handle, int(our_level, c_long_long), int(which, c_int), inv_mat%v)
      else if (mat_type == MATSHELL) then
         call MatShellGetContext(inv_mat, mat_ctx, ierr)
         if (associated(mat_ctx%real_roots)) then        ! Newton basis
            ierr_c = pflare_b200_set_poly(handle, int(our_level, c_int), int(which, c_int), int(inverse_type, c_int), &
                        int(size(mat_ctx%real_roots), c_int), mat_ctx%real_roots, mat_ctx%imag_roots, &
                        merge(1_c_int, 0_c_int, diag_scale))
         else                                            ! power / Arnoldi / Neumann coefficients
            dummy_im = 0d0
            ierr_c = pflare_b200_set_poly(handle, int(our_level, c_int), int(which, c_int), int(inverse_type, c_int), &
                        int(size(mat_ctx%coefficients), c_int), mat_ctx%coefficients, dummy_im, &
                        merge(1_c_int, 0_c_int, diag_scale))
         end if
         if (ierr_c /= 0) call MPI_Abort(MPI_COMM_WORLD, MPI_ERR_OTHER, ierr)
      else
         call pflare_b200_upload_mat_c(handle, int(our_level, c_long_long), int(which, c_int), inv_mat%v)
      end if
   end subroutine upload_inverse_b200

   ! Upload hook: call at the end of setup_air_pcmg, after the temp vecs (src/AIR_MG_Setup.F90:1211)
   subroutine upload_air_data_b200(air_data, amat)
      type(air_multigrid_data), intent(inout) :: air_data
      type(tMat), intent(in)                  :: amat

      integer :: our_level, no_levels, ierr_abort
      integer(c_int) :: ierr_c
      logical :: any_c
      type(tMat) :: level_mat

      no_levels = air_data%no_levels
      ! The library keeps PETSc's ownership on every level (x_c of level l IS x of level l+1 on the same rank).  Levels
      ! that the reference repartitions onto fewer ranks (-pc_air_processor_agglom, default .TRUE.,
      ! src/AIR_Data_Type.F90:64, src/AIR_MG_Setup.F90:645-907) break that: run with -pc_air_processor_agglom 0 and let
      ! the library agglomerate its coarse levels itself (option "agg_rows"); finalize_setup fails loudly otherwise
      ! ("level l+1 has X rows but level l has Y C points").
      if (air_data%options%processor_agglom) then
         print *, "pflare_b200: -pc_air_processor_agglom must be 0 (the library agglomerates coarse levels itself)"
         call MPI_Abort(MPI_COMM_WORLD, MPI_ERR_OTHER, ierr_abort)
      end if
      call pflare_b200_create_c(air_data%b200_handle, air_data%b200_aux, amat%v, int(no_levels, c_long_long))
      if (air_data%options%full_smoothing_up_and_down) then
         ierr_c = pflare_b200_set_option(air_data%b200_handle, "full_smoothing_up_and_down"//c_null_char, 1d0)
      end if
      ! -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it N: the coarse KSP of the PCMG (src/AIR_MG_Setup.F90:1094-1102) then runs
      ! N Richardson sweeps around mg_coarse_shell_apply; KSPPREONLY (the default) is one application
      call pflare_b200_coarse_its_c(air_data%pcmg%v, ierr_c)       ! PCMGGetCoarseSolve -> KSPGetType / KSPGetTolerances (shim C file)
      if (ierr_c > 1) then
         ierr_c = pflare_b200_set_option(air_data%b200_handle, "mg_coarse_ksp_max_it"//c_null_char, real(ierr_c, c_double))
      end if
      do our_level = 1, no_levels - 1
         ! the level operator fixes the row ownership of the level: level 1 = amat, else coarse_matrix(our_level)
         level_mat = amat
         if (our_level > 1) level_mat = air_data%coarse_matrix(our_level)
         call pflare_b200_set_level_c(air_data%b200_handle, int(our_level, c_long_long), level_mat%v, &
                  air_data%IS_fine_index(our_level)%v, air_data%IS_coarse_index(our_level)%v, &
                  int(air_data%smooth_order_levels(our_level)%array, c_int), &
                  int(size(air_data%smooth_order_levels(our_level)%array), c_int))
         if (air_data%options%full_smoothing_up_and_down) then
            ! the smoother inverts the whole level matrix (src/AIR_MG_Setup.F90:1014-1024): hand over coarse_matrix(level)
            call pflare_b200_upload_mat_c(air_data%b200_handle, int(our_level, c_long_long), B200_COARSE, level_mat%v)
         else
            call pflare_b200_upload_mat_c(air_data%b200_handle, int(our_level, c_long_long), B200_AFF, air_data%A_ff(our_level)%v)
            call pflare_b200_upload_mat_c(air_data%b200_handle, int(our_level, c_long_long), B200_AFC, air_data%A_fc(our_level)%v)
         end if
         call pflare_b200_upload_mat_c(air_data%b200_handle, int(our_level, c_long_long), B200_R, air_data%restrictors(our_level)%v)
         call pflare_b200_upload_mat_c(air_data%b200_handle, int(our_level, c_long_long), B200_P, air_data%prolongators(our_level)%v)
         call upload_inverse_b200(air_data%b200_handle, our_level, B200_INV_AFF, air_data%inv_A_ff(our_level), &
                  air_data%options%inverse_type, air_data%options%diag_scale_polys)
         any_c = any(air_data%smooth_order_levels(our_level)%array < 0) .AND. .NOT. air_data%options%full_smoothing_up_and_down
         if (any_c) then
            call pflare_b200_upload_mat_c(air_data%b200_handle, int(our_level, c_long_long), B200_ACF, air_data%A_cf(our_level)%v)
            call pflare_b200_upload_mat_c(air_data%b200_handle, int(our_level, c_long_long), B200_ACC, air_data%A_cc(our_level)%v)
            call upload_inverse_b200(air_data%b200_handle, our_level, B200_INV_ACC, air_data%inv_A_cc(our_level), &
                     air_data%options%c_inverse_type, air_data%options%diag_scale_polys)
         end if
      end do
      ! coarsest level: coarse_matrix(no_levels) (needed by a matrix-free coarse solver) + inv_A_ff(no_levels)
      call pflare_b200_set_level_c(air_data%b200_handle, int(no_levels, c_long_long), air_data%coarse_matrix(no_levels)%v, &
               PETSC_NULL_IS%v, PETSC_NULL_IS%v, [0_c_int], 0_c_int)
      call pflare_b200_upload_mat_c(air_data%b200_handle, int(no_levels, c_long_long), B200_COARSE, air_data%coarse_matrix(no_levels)%v)
      call upload_inverse_b200(air_data%b200_handle, no_levels, B200_INV_AFF, air_data%inv_A_ff(no_levels), &
               air_data%options%coarsest_inverse_type, air_data%options%coarsest_diag_scale_polys)
      call pflare_b200_finalize_c(air_data%b200_handle)
   end subroutine upload_air_data_b200

   ! Replaces `call PCApply(pcmg, x, y, ierr)` in PCApply_AIR_Shell (src/PCAIR_Shell.F90:170-188)
   subroutine apply_air_b200(air_data, x, y, ierr)
      type(air_multigrid_data), intent(inout) :: air_data
      type(tVec), intent(in)    :: x
      type(tVec), intent(inout) :: y
      PetscErrorCode, intent(out) :: ierr
      call pflare_b200_apply_c(air_data%b200_handle, air_data%b200_aux, x%v, y%v)
      ierr = 0          ! Fortran PETSc callbacks must set ierr (src/FC_Smooth.F90:492-493)
   end subroutine apply_air_b200

   ! Call from reset_air_data (src/AIR_Data_Type_Routines.F90:105)
   subroutine destroy_air_b200(air_data)
      type(air_multigrid_data), intent(inout) :: air_data
      if (c_associated(air_data%b200_handle)) call pflare_b200_destroy_c(air_data%b200_handle, air_data%b200_aux)
      air_data%b200_handle = c_null_ptr
      air_data%b200_aux = c_null_ptr
   end subroutine destroy_air_b200

end module pflare_b200_mod
