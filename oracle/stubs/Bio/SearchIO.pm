# oracle/stubs/Bio/SearchIO.pm -- TEST INFRASTRUCTURE ONLY: lib/SeqToolBox/HMMER/Parser.pm:42 imports BioPerl's
# Bio::SearchIO without using it on the paths exercised here; BioPerl is not installed in this image.
package Bio::SearchIO;
1;
