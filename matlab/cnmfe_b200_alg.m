function a = cnmfe_b200_alg(name)
%% options.spatial_algorithm -> code of the library (update_spatial_parallel.m:202-212; anything else falls to nnls there)
switch lower(name)
    case 'hals';        a = 0;
    case 'hals_thresh'; a = 1;
    case 'lars';        a = 3;
    otherwise;          a = 2;     % 'nnls' and the reference's `otherwise` branch
end
end
